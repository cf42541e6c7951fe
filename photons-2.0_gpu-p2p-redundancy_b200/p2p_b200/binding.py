"""ctypes binding of lib/libp2p_b200.so (native C-ABI, include/p2p_b200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.environ.get("P2P_B200_LIB") or os.path.join(LIB_DIR, "libp2p_b200.so")     # the override serves kernel A/B builds (tools/README.md)

KERNEL_AUTO, KERNEL_SCALAR, KERNEL_PACKED = 0, 1, 2


class P2PError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"p2p_b200 error {code}: {msg}")
        self.code = code


def build_library(force=False):
    """Compile the sm_100a libraries in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", PKG_DIR, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", PKG_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise P2PError(-4, f"{LIB_PATH} is not built (run __graft_entry__.build() or make -C {PKG_DIR}); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.p2p_last_error.restype = C.c_char_p
    L.p2p_device_particles.restype = C.c_void_p
    L.p2p_device_acc.restype = C.c_void_p
    L.p2p_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    for name in ("p2p_destroy", "p2p_clear_ghosts", "p2p_clear_tasks", "p2p_build_csr", "p2p_compute", "p2p_zero_acc",
                 "p2p_synchronize"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.p2p_set_physics.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.p2p_set_kernel_variant.argtypes = [C.c_void_p, C.c_int]
    L.p2p_set_box.argtypes = [C.c_void_p, _dp, C.c_double]
    L.p2p_set_tuning.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.p2p_set_far_threshold.argtypes = [C.c_void_p, C.c_double]
    L.p2p_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.p2p_upload_particles.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64]
    L.p2p_upload_leaves.argtypes = [C.c_void_p, _ip, _ip, C.c_int]
    L.p2p_append_ghosts.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _ip, _ip, C.c_int, _ip]
    L.p2p_append_ghosts_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, _ip, _ip, C.c_int, _ip]
    L.p2p_append_tasks.argtypes = [C.c_void_p, _ip, _ip, C.c_int64, C.c_int]
    L.p2p_append_tasks_interleaved.argtypes = [C.c_void_p, _ip, C.c_int64, C.c_int]
    L.p2p_download_acc.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int]
    L.p2p_counts.argtypes = [C.c_void_p, _lp, _lp]
    L.p2p_accumulated_counts.argtypes = [C.c_void_p, _lp, _lp]
    L.p2p_download_csr.argtypes = [C.c_void_p, _lp, _ip]
    L.p2p_download_csr_class.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), _ip]
    L.p2p_last_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.p2p_step_host.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _ip, _ip, C.c_int, _ip, _ip, C.c_int64, _dp,
                                C.c_int64, C.c_int]
    L.p2p_step_host_chunked.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _ip, _ip, C.c_int, _dp, C.c_int64, C.c_int64, _ip,
                                        _ip, C.c_int, _ip, _ip, _lp, C.c_int, _dp, C.c_int64, C.c_int]
    L.p2p_device_particles.argtypes = [C.c_void_p]
    L.p2p_tree_build.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int]
    L.p2p_tree_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip, _dp, _dp]
    L.p2p_tree_info.argtypes = [C.c_void_p, _ip, _ip, _ip, C.POINTER(C.c_float), C.POINTER(C.c_float), _lp, _dp]
    L.p2p_step_device.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
    L.p2p_tree_download.argtypes = [C.c_void_p, _lp, _dp, _ip, _ip, _dp, _dp, _ip, _ip, _dp, _dp, _dp]
    L.p2p_tree_walk.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp, _dp]
    L.p2p_csr_duplicates.argtypes = [C.c_void_p, _lp]
    L.p2p_download_acc_original.argtypes = [C.c_void_p, _dp]
    L.p2p_tree_set_option.argtypes = [C.c_void_p, C.c_int]
    L.p2p_route_load.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int]
    L.p2p_route_partition.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
    L.p2p_route_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.p2p_route_import.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.p2p_tree_build_resident.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int]
    L.p2p_download_index.argtypes = [C.c_void_p, _lp]
    L.p2p_resident_load.argtypes = [C.c_void_p, _dp, C.c_int64, _dp, C.c_int64, C.c_int64, C.c_int64]
    L.p2p_resident_build.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int]
    L.p2p_resident_partition.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
    L.p2p_resident_export.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    L.p2p_resident_import.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64]
    L.p2p_resident_forces.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.p2p_resident_kick.argtypes = [C.c_void_p, C.c_double]
    L.p2p_resident_drift.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.p2p_resident_download.argtypes = [C.c_void_p, _dp, _dp, _lp]
    L.p2p_midfield_enable.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.p2p_midfield_compute.argtypes = [C.c_void_p, _lp]
    L.p2p_midfield_multipoles.argtypes = [C.c_void_p, C.c_void_p]
    L.p2p_midfield_download.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.POINTER(C.c_float)]
    L.p2p_device_acc.argtypes = [C.c_void_p]
    L.p2p_resident_count.argtypes = [C.c_void_p, _lp]
    L.p2p_forces_local.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_int]
    L.p2p_set_chunk_tasks.argtypes = [C.c_void_p, C.c_int64]
    L.p2p_set_chunk_pipeline.argtypes = [C.c_void_p, C.c_int]
    L.p2p_step_timings.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 4 + [_ip]
    L.p2p_swap_lists.argtypes = [C.c_void_p]
    L.p2p_set_force_blocks.argtypes = [C.c_void_p, C.c_int]
    L.p2p_reserve_ghosts.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.p2p_tree_walk_range.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_int]
    L.p2p_set_rank.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.p2p_topology_stride.argtypes = [C.c_int, C.c_int, _lp]
    L.p2p_tree_export_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.p2p_tree_walk_peers_packed.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_int, _ip, _ip,
                                             C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.p2p_halo_plan_need.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _ip, C.c_int, C.c_int, C.c_void_p, _lp]
    L.p2p_halo_plan_give.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _lp]
    L.p2p_halo_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.p2p_halo_set_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.p2p_midfield_compute_peers_packed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, _lp]
    _lib = L
    return L


def device_count():
    return int(load_library().p2p_device_count())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class P2PContext:
    """One device context: persistent particles, leaves, ghost leaves, CSR task list, accelerations."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        self._chk(self._L.p2p_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        self.npart = 0
        self.nleaf = 0
        self.stream_ptr = 0                          # externally owned stream, 0 = the context's own

    def _chk(self, rc):
        if rc != 0:
            raise P2PError(rc, self._L.p2p_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.p2p_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_physics(self, mass, eps, rs):
        self._chk(self._L.p2p_set_physics(self._h, float(mass), float(eps), float(rs)))

    def set_kernel_variant(self, v):
        self._chk(self._L.p2p_set_kernel_variant(self._h, int(v)))

    def set_box(self, origin, extent):
        o = _f64(origin)
        assert o.shape == (3,)
        self._chk(self._L.p2p_set_box(self._h, o.ctypes.data_as(_dp), float(extent)))

    def set_tuning(self, targets_per_pass=0, sources_per_lane=0, min_blocks=0):
        self._chk(self._L.p2p_set_tuning(self._h, int(targets_per_pass), int(sources_per_lane), int(min_blocks)))

    def set_far_threshold(self, u_far=-1.0):
        """near / far split of the list in u = r / 2 r_s (< 0: default, 0: every source leaf through the full kernel body)"""
        self._chk(self._L.p2p_set_far_threshold(self._h, float(u_far)))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self._L.p2p_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))
        self.stream_ptr = int(cuda_stream_ptr or 0)

    def upload_particles(self, pos):
        pos = _f64(pos)
        assert pos.ndim == 2 and pos.shape[1] >= 3
        self._chk(self._L.p2p_upload_particles(self._h, pos.ctypes.data_as(_dp), pos.shape[1], pos.shape[0]))
        self.npart = pos.shape[0]

    def upload_leaves(self, leaf_npart, leaf_ipart):
        n, i = _i32(leaf_npart), _i32(leaf_ipart)
        self._chk(self._L.p2p_upload_leaves(self._h, n.ctypes.data_as(_ip), i.ctypes.data_as(_ip), len(n)))
        self.nleaf = len(n)

    def append_ghosts(self, pos, start, count):
        pos = _f64(pos).reshape(-1, 3) if np.size(pos) else np.zeros((0, 3))
        s, c = _i32(start), _i32(count)
        first = C.c_int()
        self._chk(self._L.p2p_append_ghosts(self._h, pos.ctypes.data_as(_dp), 3, pos.shape[0], s.ctypes.data_as(_ip),
                                            c.ctypes.data_as(_ip), len(s), C.byref(first)))
        return first.value

    def append_ghosts_device(self, dev_ptr, nbody, start, count):
        s, c = _i32(start), _i32(count)
        first = C.c_int()
        self._chk(self._L.p2p_append_ghosts_device(self._h, C.c_void_p(dev_ptr), int(nbody), s.ctypes.data_as(_ip),
                                                   c.ctypes.data_as(_ip), len(s), C.byref(first)))
        return first.value

    def clear_ghosts(self):
        self._chk(self._L.p2p_clear_ghosts(self._h))

    def clear_tasks(self):
        self._chk(self._L.p2p_clear_tasks(self._h))

    def append_tasks(self, tt, ts, source_offset=0):
        tt, ts = _i32(tt), _i32(ts)
        assert len(tt) == len(ts)
        self._chk(self._L.p2p_append_tasks(self._h, tt.ctypes.data_as(_ip), ts.ctypes.data_as(_ip), len(tt),
                                           int(source_offset)))

    def append_tasks_interleaved(self, pairs, source_offset=0):
        pairs = _i32(pairs).reshape(-1, 2)
        self._chk(self._L.p2p_append_tasks_interleaved(self._h, pairs.ctypes.data_as(_ip), pairs.shape[0],
                                                       int(source_offset)))

    def build_csr(self):
        self._chk(self._L.p2p_build_csr(self._h))

    def compute(self):
        self._chk(self._L.p2p_compute(self._h))

    def zero_acc(self):
        self._chk(self._L.p2p_zero_acc(self._h))

    def synchronize(self):
        self._chk(self._L.p2p_synchronize(self._h))

    def download_acc(self, out=None, accumulate=False):
        if out is None:
            out = np.zeros((self.npart, 3))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape[0] == self.npart
        self._chk(self._L.p2p_download_acc(self._h, out.ctypes.data_as(_dp), out.shape[1], 1 if accumulate else 0))
        return out

    def counts(self):
        nt, npairs = C.c_int64(), C.c_int64()
        self._chk(self._L.p2p_counts(self._h, C.byref(nt), C.byref(npairs)))
        return nt.value, npairs.value

    def accumulated_counts(self):
        nt, npairs = C.c_int64(), C.c_int64()
        self._chk(self._L.p2p_accumulated_counts(self._h, C.byref(nt), C.byref(npairs)))
        return nt.value, npairs.value

    def download_csr(self, raw=False):
        """(row_ptr, col) of the packed list.  On the device a row holds its far source leaves first, then its near ones
        (download_csr_class), each class in ascending order; raw=False returns the canonical form (whole rows ascending)."""
        nt, _ = self.counts()
        row = np.zeros(self.nleaf + 1, np.int64)
        col = np.zeros(max(nt, 1), np.int32)
        self._chk(self._L.p2p_download_csr(self._h, row.ctypes.data_as(_lp), col.ctypes.data_as(_ip)))
        col = col[:nt]
        if not raw and nt:
            rid = np.repeat(np.arange(self.nleaf), np.diff(row))
            col = col[np.lexsort((col, rid))]
        return row, col

    def download_csr_class(self):
        """(is_far per CSR column, number of far columns per row -- they come first): the split the force kernel uses"""
        nt, _ = self.counts()
        far = np.zeros(max(nt, 1), np.uint8)
        nfar = np.zeros(max(self.nleaf, 1), np.int32)
        self._chk(self._L.p2p_download_csr_class(self._h, far.ctypes.data_as(C.POINTER(C.c_ubyte)), nfar.ctypes.data_as(_ip)))
        return far[:nt], nfar[:self.nleaf]

    def last_timings(self):
        a, b = C.c_float(), C.c_float()
        self._chk(self._L.p2p_last_timings(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def step_host(self, pos, leaf_npart, leaf_ipart, tt, ts, acc=None, accumulate=False):
        """Host buffers in, host accelerations out (H2D + CSR + kernel + D2H)."""
        pos = _f64(pos)
        n, i, tt, ts = _i32(leaf_npart), _i32(leaf_ipart), _i32(tt), _i32(ts)
        if acc is None:
            acc = np.zeros((pos.shape[0], 3))
        self._chk(self._L.p2p_step_host(self._h, pos.ctypes.data_as(_dp), pos.shape[1], pos.shape[0],
                                        n.ctypes.data_as(_ip), i.ctypes.data_as(_ip), len(n), tt.ctypes.data_as(_ip),
                                        ts.ctypes.data_as(_ip), len(tt), acc.ctypes.data_as(_dp), acc.shape[1],
                                        1 if accumulate else 0))
        self.npart, self.nleaf = pos.shape[0], len(n)
        return acc

    def step_host_chunked(self, pos, leaf_npart, leaf_ipart, tt, ts, chunk_off, ghost_pos=None, ghost_start=None,
                          ghost_count=None, acc=None, accumulate=False):
        """Chunk-pipelined host step: H2D + packing of chunk g+1 overlap the kernel of chunk g.
        Arrays are used as given (no copies): pass pinned, contiguous arrays of the right dtype."""
        assert pos.dtype == np.float64 and tt.dtype == np.int32 and ts.dtype == np.int32
        off = np.ascontiguousarray(chunk_off, np.int64)
        n, i = _i32(leaf_npart), _i32(leaf_ipart)
        if ghost_pos is None or len(ghost_start) == 0:
            gp, gs, gc = np.zeros((0, 3)), np.zeros(0, np.int32), np.zeros(0, np.int32)
        else:
            gp, gs, gc = _f64(ghost_pos).reshape(-1, 3), _i32(ghost_start), _i32(ghost_count)
        if acc is None:
            acc = np.zeros((pos.shape[0], 3))
        self._chk(self._L.p2p_step_host_chunked(
            self._h, pos.ctypes.data_as(_dp), pos.shape[1], pos.shape[0], n.ctypes.data_as(_ip), i.ctypes.data_as(_ip), len(n),
            gp.ctypes.data_as(_dp), 3, gp.shape[0], gs.ctypes.data_as(_ip), gc.ctypes.data_as(_ip), len(gs),
            tt.ctypes.data_as(_ip), ts.ctypes.data_as(_ip), off.ctypes.data_as(_lp), len(off) - 1, acc.ctypes.data_as(_dp),
            acc.shape[1], 1 if accumulate else 0))
        self.npart, self.nleaf = pos.shape[0], len(n)
        return acc

    # ---- device-resident tree build / dual-tree walk
    def tree_build(self, pos, maxleaf, bdl, bdr, direct_start=0):
        """build_localtree on the device from positions in the caller's order (no copy: pass a contiguous float64 array)."""
        pos = _f64(pos)
        bl, br = np.ascontiguousarray(bdl, np.float64), np.ascontiguousarray(bdr, np.float64)
        self._chk(self._L.p2p_tree_build(self._h, pos.ctypes.data_as(_dp), pos.shape[1], pos.shape[0], int(maxleaf),
                                         bl.ctypes.data_as(_dp), br.ctypes.data_as(_dp), int(direct_start)))
        self.npart = pos.shape[0]
        self.nleaf = self.tree_info()["nleaf"]
        self._tree_maxleaf = int(maxleaf)

    def tree_upload(self, tree):
        """Boxes and sons of a host-built tree (host.LocalTree, or any object with the same attributes) for tree_walk."""
        nl, nn = int(tree.nleaf), int(tree.nnode)
        lc, lw = _f64(tree.leaf_center[:nl]), _f64(tree.leaf_width[:nl])
        son, nc, nw = _i32(tree.node_son[:nn]), _f64(tree.node_center[:nn]), _f64(tree.node_width[:nn])
        self._chk(self._L.p2p_tree_upload(self._h, int(tree.maxleaf), nl, nn, int(tree.first_leaf), int(tree.first_node),
                                          lc.ctypes.data_as(_dp), lw.ctypes.data_as(_dp), son.ctypes.data_as(_ip),
                                          nc.ctypes.data_as(_dp), nw.ctypes.data_as(_dp)))

    def tree_info(self):
        nl, nn, nv = C.c_int(), C.c_int(), C.c_int()
        mb, mw, it, wd = C.c_float(), C.c_float(), C.c_int64(), C.c_double()
        self._chk(self._L.p2p_tree_info(self._h, C.byref(nl), C.byref(nn), C.byref(nv), C.byref(mb), C.byref(mw), C.byref(it),
                                        C.byref(wd)))
        return dict(nleaf=nl.value, nnode=nn.value, nlevel=nv.value, ms_build=mb.value, ms_walk=mw.value, walk_items=it.value,
                    max_leaf_width=wd.value)

    def tree_download(self):
        """The device-built tree in the reference's layout (same keys as oracle.Tree / host.LocalTree)."""
        info = self.tree_info()
        nl, nn, n = info["nleaf"], info["nnode"], self.npart
        out = dict(nleaf=nl, nnode=nn, perm=np.zeros(n, np.int64), pos=np.zeros((n, 3)),
                   leaf_npart=np.zeros(nl, np.int32), leaf_ipart=np.zeros(nl, np.int32), leaf_center=np.zeros((nl, 3)),
                   leaf_width=np.zeros((nl, 3)), node_npart=np.zeros(nn, np.int32), node_son=np.zeros((nn, 2), np.int32),
                   node_split=np.zeros(nn), node_center=np.zeros((nn, 3)), node_width=np.zeros((nn, 3)))
        self._chk(self._L.p2p_tree_download(
            self._h, out["perm"].ctypes.data_as(_lp), out["pos"].ctypes.data_as(_dp), out["leaf_npart"].ctypes.data_as(_ip),
            out["leaf_ipart"].ctypes.data_as(_ip), out["leaf_center"].ctypes.data_as(_dp), out["leaf_width"].ctypes.data_as(_dp),
            out["node_npart"].ctypes.data_as(_ip), out["node_son"].ctypes.data_as(_ip), out["node_split"].ctypes.data_as(_dp),
            out["node_center"].ctypes.data_as(_dp), out["node_width"].ctypes.data_as(_dp)))
        return out

    def tree_walk(self, theta, rcut, period=0.0, tcenter=None, twidth=None):
        """Dual-tree walk on the device (local list, plus the 26 periodic images when period > 0); appends tasks."""
        tc = np.ascontiguousarray(tcenter if tcenter is not None else np.zeros(3), np.float64)
        tw = np.ascontiguousarray(twidth if twidth is not None else np.zeros(3), np.float64)
        self._chk(self._L.p2p_tree_walk(self._h, float(theta), float(rcut), float(period), tc.ctypes.data_as(_dp),
                                        tw.ctypes.data_as(_dp)))

    def csr_duplicates(self):
        d = C.c_int64()
        self._chk(self._L.p2p_csr_duplicates(self._h, C.byref(d)))
        return d.value

    def download_acc_original(self, out=None):
        if out is None:
            out = np.zeros((self.npart, 3))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (self.npart, 3)
        self._chk(self._L.p2p_download_acc_original(self._h, out.ctypes.data_as(_dp)))
        return out

    def step_device(self, pos, maxleaf, bdl, bdr, theta, rcut, period=0.0, direct_start=0, acc=None):
        """Positions (caller's order) in, accelerations (same order) out; everything in between on the device.
        No copies: pass contiguous float64 arrays (pinned for full PCIe rate)."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous and pos.ndim == 2 and pos.shape[1] >= 3
        if acc is None:
            acc = np.zeros((pos.shape[0], 3))
        assert acc.dtype == np.float64 and acc.flags.c_contiguous and acc.shape == (pos.shape[0], 3)
        bl, br = np.ascontiguousarray(bdl, np.float64), np.ascontiguousarray(bdr, np.float64)
        self._chk(self._L.p2p_step_device(self._h, pos.ctypes.data_as(_dp), pos.shape[1], pos.shape[0], int(maxleaf),
                                          bl.ctypes.data_as(_dp), br.ctypes.data_as(_dp), int(direct_start), float(theta),
                                          float(rcut), float(period), acc.ctypes.data_as(_dp)))
        self.npart = pos.shape[0]
        self.nleaf = self.tree_info()["nleaf"]
        return acc

    # ---- particle routing on the device (multi-rank)
    def route_load(self, pos, first_index, append=False):
        """host slab -> resident device arrays (global ids first_index ...); append=True adds a further piece of the slab"""
        pos = _f64(pos)
        self._chk(self._L.p2p_route_load(self._h, pos.ctypes.data_as(_dp), pos.shape[1] if pos.ndim == 2 else 3, pos.shape[0],
                                         int(first_index), 1 if append else 0))

    def route_partition(self, nproc, split):
        sp = np.ascontiguousarray(split, np.float64)
        send = np.zeros(nproc, np.int32)
        self._chk(self._L.p2p_route_partition(self._h, int(nproc), sp.ctypes.data_as(_dp), send.ctypes.data_as(_ip)))
        return send

    def route_export(self, d_x, d_y, d_z, d_idx):
        self._chk(self._L.p2p_route_export(self._h, d_x, d_y, d_z, d_idx))

    def route_import(self, d_x, d_y, d_z, d_idx, n):
        self._chk(self._L.p2p_route_import(self._h, d_x, d_y, d_z, d_idx, int(n)))
        self._resident = int(n)

    def tree_build_resident(self, maxleaf, bdl, bdr, direct_start=0):
        bl, br = np.ascontiguousarray(bdl, np.float64), np.ascontiguousarray(bdr, np.float64)
        self._chk(self._L.p2p_tree_build_resident(self._h, int(maxleaf), bl.ctypes.data_as(_dp), br.ctypes.data_as(_dp), int(direct_start)))
        info = self.tree_info()
        self.nleaf = info["nleaf"]
        self.npart = self.resident_count()

    def download_index(self, out=None):
        idx = np.zeros(self.npart, np.int64) if out is None else out
        assert idx.dtype == np.int64 and idx.flags.c_contiguous and idx.shape == (self.npart,)
        self._chk(self._L.p2p_download_index(self._h, idx.ctypes.data_as(_lp)))
        return idx

    # ---- device-resident stepping
    def resident_load(self, pos, vel=None, first_id=0):
        pos = _f64(pos)
        v = _f64(vel) if vel is not None else None
        self._chk(self._L.p2p_resident_load(self._h, pos.ctypes.data_as(_dp), pos.shape[1], v.ctypes.data_as(_dp) if v is not None else None,
                                            v.shape[1] if v is not None else 3, pos.shape[0], int(first_id)))
        self.npart = pos.shape[0]

    def resident_build(self, maxleaf, bdl, bdr, direct_start=0):
        bl, br = np.ascontiguousarray(bdl, np.float64), np.ascontiguousarray(bdr, np.float64)
        self._chk(self._L.p2p_resident_build(self._h, int(maxleaf), bl.ctypes.data_as(_dp), br.ctypes.data_as(_dp), int(direct_start)))
        self.nleaf = self.tree_info()["nleaf"]
        self.npart = self.resident_count()

    def resident_partition(self, nproc, split):
        split = _f64(split)
        send = np.zeros(nproc, np.int32)
        self._chk(self._L.p2p_resident_partition(self._h, int(nproc), split.ctypes.data_as(_dp), send.ctypes.data_as(_ip)))
        return send

    def resident_export(self, d_xv, d_id):
        arr = (C.c_void_p * 6)(*[C.c_void_p(int(p)) for p in d_xv])
        self._chk(self._L.p2p_resident_export(self._h, arr, C.c_void_p(int(d_id))))

    def resident_import(self, d_xv, d_id, n):
        arr = (C.c_void_p * 6)(*[C.c_void_p(int(p)) for p in d_xv])
        self._chk(self._L.p2p_resident_import(self._h, arr, C.c_void_p(int(d_id)), int(n)))
        self.npart = int(n)

    def resident_forces(self, maxleaf, bdl, bdr, theta, rcut, period, direct_start=0):
        bl, br = np.ascontiguousarray(bdl, np.float64), np.ascontiguousarray(bdr, np.float64)
        self._chk(self._L.p2p_resident_forces(self._h, int(maxleaf), bl.ctypes.data_as(_dp), br.ctypes.data_as(_dp), int(direct_start),
                                              float(theta), float(rcut), float(period)))
        self.nleaf = self.tree_info()["nleaf"]

    def resident_kick(self, dkh):
        self._chk(self._L.p2p_resident_kick(self._h, float(dkh)))

    def resident_drift(self, dd, period):
        self._chk(self._L.p2p_resident_drift(self._h, float(dd), float(period)))

    def resident_download(self):
        n = self.npart
        pos, vel, idx = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n, np.int64)
        self._chk(self._L.p2p_resident_download(self._h, pos.ctypes.data_as(_dp), vel.ctypes.data_as(_dp), idx.ctypes.data_as(_lp)))
        return pos, vel, idx

    # ---- mid-field (M2L lists from the walk; P2M / M2M / M2L / L2L / L2P kernels)
    def midfield_enable(self, on=True, literal_d6=False):
        self._chk(self._L.p2p_midfield_enable(self._h, 1 if on else 0, 1 if literal_d6 else 0))

    def midfield_compute(self):
        n = C.c_int64()
        self._chk(self._L.p2p_midfield_compute(self._h, C.byref(n)))
        return n.value

    def midfield_multipoles(self, d_M):
        self._chk(self._L.p2p_midfield_multipoles(self._h, d_M))

    def midfield_download(self):
        info = self.tree_info()
        nl, nn = info["nleaf"], info["nnode"]
        out = dict(leaf_M=np.zeros((nl, 20)), node_M=np.zeros((nn, 20)), leaf_L=np.zeros((nl, 20)), node_L=np.zeros((nn, 20)))
        ms = C.c_float()
        self._chk(self._L.p2p_midfield_download(self._h, out["leaf_M"].ctypes.data_as(_dp), out["node_M"].ctypes.data_as(_dp),
                                                out["leaf_L"].ctypes.data_as(_dp), out["node_L"].ctypes.data_as(_dp), C.byref(ms)))
        out["ms"] = ms.value
        return out

    # ---- multi-rank device path (device pointers, e.g. tensor.data_ptr())
    # ---- multi-rank exchange: packed topology, local / remote phases, device-side halo plan
    def resident_count(self):
        n = C.c_int64()
        self._chk(self._L.p2p_resident_count(self._h, C.byref(n)))
        return n.value

    def forces_local(self, theta, rcut, period, tcenter, twidth, compute=True):
        """walk (own tree + own periodic images), packing and forces of the device-built tree, in target chunks"""
        tc, tw = _f64(tcenter), _f64(twidth)
        self._chk(self._L.p2p_forces_local(self._h, float(theta), float(rcut), float(period), tc.ctypes.data_as(_dp), tw.ctypes.data_as(_dp),
                                           1 if compute else 0))

    def set_chunk_tasks(self, max_tasks):
        self._chk(self._L.p2p_set_chunk_tasks(self._h, int(max_tasks)))

    def set_chunk_pipeline(self, min_chunks=0):
        """single-rank forces_local: at least min_chunks target chunks, walk + packing of the next beside the forces of the current (0: off)"""
        self._chk(self._L.p2p_set_chunk_pipeline(self._h, int(min_chunks)))

    def step_timings(self):
        v = [C.c_float() for _ in range(4)]
        n = C.c_int()
        self._chk(self._L.p2p_step_timings(self._h, *[C.byref(x) for x in v], C.byref(n)))
        return dict(build_ms=v[0].value, walk_ms=v[1].value, csr_ms=v[2].value, force_ms=v[3].value, chunks=n.value)

    def swap_lists(self):
        self._chk(self._L.p2p_swap_lists(self._h))

    def set_force_blocks(self, budget=0):
        """0: persistent force-kernel warps; k: warps retire after k x 2^17 cycles (slots for higher-priority streams)"""
        self._chk(self._L.p2p_set_force_blocks(self._h, int(budget)))

    def reserve_ghosts(self, nghostleaf, nghost):
        self._chk(self._L.p2p_reserve_ghosts(self._h, int(nghostleaf), int(nghost)))

    def tree_walk_range(self, theta, rcut, period, tcenter, twidth, leaf_lo, leaf_hi):
        tc, tw = _f64(tcenter), _f64(twidth)
        self._chk(self._L.p2p_tree_walk_range(self._h, float(theta), float(rcut), float(period), tc.ctypes.data_as(_dp), tw.ctypes.data_as(_dp),
                                              int(leaf_lo), int(leaf_hi)))

    def set_rank(self, rank, nranks):
        self._chk(self._L.p2p_set_rank(self._h, int(rank), int(nranks)))

    def topology_stride(self, nleaf_max, nnode_max):
        v = C.c_int64()
        self._chk(self._L.p2p_topology_stride(int(nleaf_max), int(nnode_max), C.byref(v)))
        return v.value

    def tree_export_packed(self, d_block, nleaf_max, nnode_max):
        self._chk(self._L.p2p_tree_export_packed(self._h, C.c_void_p(d_block), int(nleaf_max), int(nnode_max)))

    def tree_walk_peers_packed(self, theta, rcut, period, tcenter, twidth, me, peer_nleaf, peer_nnode, d_all, nleaf_max, nnode_max, include_me):
        tc, tw = _f64(tcenter), _f64(twidth)
        nl, nn = _i32(peer_nleaf), _i32(peer_nnode)
        self._chk(self._L.p2p_tree_walk_peers_packed(self._h, float(theta), float(rcut), float(period), tc.ctypes.data_as(_dp),
                                                     tw.ctypes.data_as(_dp), len(nl), int(me), nl.ctypes.data_as(_ip), nn.ctypes.data_as(_ip),
                                                     C.c_void_p(d_all), int(nleaf_max), int(nnode_max), int(include_me)))

    def halo_plan_need(self, d_topo_all, me, peer_nleaf, nleaf_max, nnode_max, d_marks):
        nl = _i32(peer_nleaf)
        out = np.zeros(len(nl), np.int64)
        self._chk(self._L.p2p_halo_plan_need(self._h, C.c_void_p(d_topo_all), len(nl), int(me), nl.ctypes.data_as(_ip), int(nleaf_max),
                                             int(nnode_max), C.c_void_p(d_marks), out.ctypes.data_as(_lp)))
        return out

    def halo_plan_give(self, d_asked, nreq):
        out = np.zeros(max(nreq, 1), np.int64)
        self._chk(self._L.p2p_halo_plan_give(self._h, C.c_void_p(d_asked), int(nreq), out.ctypes.data_as(_lp)))
        return out[:nreq]

    def halo_gather(self, d_asked, nreq, d_send):
        self._chk(self._L.p2p_halo_gather(self._h, C.c_void_p(d_asked), int(nreq), C.c_void_p(d_send)))

    def halo_set_particles(self, d_recv, nbody):
        self._chk(self._L.p2p_halo_set_particles(self._h, C.c_void_p(d_recv), int(nbody)))

    def midfield_compute_peers_packed(self, npeer, d_topo_all, nleaf_max, nnode_max, d_M_all):
        n = C.c_int64()
        self._chk(self._L.p2p_midfield_compute_peers_packed(self._h, int(npeer), C.c_void_p(d_topo_all), int(nleaf_max), int(nnode_max),
                                                            C.c_void_p(d_M_all), C.byref(n)))
        return n.value

    def tree_set_option(self, seq_sum_plain_max=-1):
        self._chk(self._L.p2p_tree_set_option(self._h, int(seq_sum_plain_max)))

    @property
    def device_particles_ptr(self):
        return self._L.p2p_device_particles(self._h)

    @property
    def device_acc_ptr(self):
        return self._L.p2p_device_acc(self._h)

