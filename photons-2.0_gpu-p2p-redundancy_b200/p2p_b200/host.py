"""ctypes binding of lib/libp2p_host.so (include/p2p_host.h): the product's own host-side tree
build, dual-tree walks, halo pruning and domain routing.  No CUDA, no oracle."""
import ctypes as C
import os

import numpy as np

from .binding import LIB_DIR, P2PError

HOST_LIB_PATH = os.path.join(LIB_DIR, "libp2p_host.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


class _TreeView(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("npart", "maxleaf", "nleaf", "nnode", "nleaf_cap", "nnode_cap", "first_leaf",
                                       "first_node")] + [
        ("leaf_npart", _ip), ("leaf_ipart", _ip), ("leaf_center", _dp), ("leaf_width", _dp), ("node_npart", _ip),
        ("node_son", _ip), ("node_split", _dp), ("node_center", _dp), ("node_width", _dp)]


class _Image(C.Structure):
    _fields_ = [("nnode", C.c_int), ("nbody", C.c_int), ("npart", _ip), ("son", _ip), ("center", _dp), ("width", _dp),
                ("body", _dp)]


_lib = None


def load_host_library():
    global _lib
    if _lib is None:
        if not os.path.isfile(HOST_LIB_PATH):
            raise P2PError(-4, f"{HOST_LIB_PATH} is not built (run __graft_entry__.build())")
        L = C.CDLL(HOST_LIB_PATH)
        L.p2p_build_localtree.argtypes = [C.POINTER(C.c_void_p), _dp, C.c_int64, _lp, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                          C.c_int]
        L.p2p_tree_free.argtypes = [C.c_void_p]
        L.p2p_tree_get.argtypes = [C.c_void_p, C.POINTER(_TreeView)]
        L.p2p_walk_task_p2p.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.POINTER(_ip), C.POINTER(_ip),
                                        C.POINTER(C.c_int64)]
        L.p2p_walk_plan_create.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
        L.p2p_walk_plan_nchunks.argtypes = [C.c_void_p]
        L.p2p_walk_plan_rows.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        L.p2p_walk_plan_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_ip), C.POINTER(_ip), C.POINTER(C.c_int64)]
        L.p2p_walk_plan_run_into.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, C.c_int64, C.POINTER(C.c_int64)]
        L.p2p_walk_plan_free.argtypes = [C.c_void_p]
        L.p2p_prepare_sendtree.argtypes = [C.c_void_p, _dp, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double,
                                           C.POINTER(_Image)]
        L.p2p_image_free.argtypes = [C.POINTER(_Image)]
        L.p2p_walk_task_p2p_ext.argtypes = [C.c_void_p, C.POINTER(_Image), C.c_double, C.c_double, C.c_int,
                                            C.POINTER(_ip), C.POINTER(_ip), C.POINTER(C.c_int64)]
        L.p2p_host_free.argtypes = [C.c_void_p]
        L.p2p_domain_setup.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, _ip]
        L.p2p_domain_relax.argtypes = [C.c_int, C.c_double, _dp, _dp]
        L.p2p_domain_boxes.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, _ip]
        L.p2p_domain_route.argtypes = [C.c_int, _dp, _dp, C.c_int64, _lp, C.c_int64, _ip]
        _lib = L
    return _lib


def _take(ptr, n, dtype):
    """Copy n items out of a library-owned array and free it."""
    L = load_host_library()
    if n == 0:
        out = np.empty(0, dtype)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    L.p2p_host_free(ptr)
    return out


def derived_params(box, nside, npart_total):
    """splitRadius, cutoffRadius, SoftenScale as 1_Indexing/src/initial.c:324-346 derives them."""
    rs = 1.25 * (box / float(nside))
    eps = 0.03 * box / pow(float(npart_total), 0.3333333)
    return rs, 4.5 * rs, eps


class Image:
    """Pruned halo image of a tree (RemoteNode/RemoteBody content)."""

    def __init__(self, npart, son, center, width, body):
        self.npart = np.ascontiguousarray(npart, np.int32)
        self.son = np.ascontiguousarray(son, np.int32).reshape(-1, 2)
        self.center = np.ascontiguousarray(center, np.float64).reshape(-1, 3)
        self.width = np.ascontiguousarray(width, np.float64).reshape(-1, 3)
        self.body = np.ascontiguousarray(body, np.float64).reshape(-1, 3)

    def _c(self):
        return _Image(len(self.npart), len(self.body), self.npart.ctypes.data_as(_ip), self.son.ctypes.data_as(_ip),
                      self.center.ctypes.data_as(_dp), self.width.ctypes.data_as(_dp), self.body.ctypes.data_as(_dp))

    def leaves(self, maxleaf):
        """(node index, first body, count) of the image's leaves."""
        idx = np.nonzero(self.npart <= maxleaf)[0].astype(np.int32)
        return idx, self.son[idx, 0].copy(), self.npart[idx].copy()


class LocalTree:
    """The local kd-tree; `pos` (N,3) float64 is copied and permuted, `perm` maps tree order -> input order."""

    def __init__(self, pos, maxleaf, bdl, bdr, direct_start=0, nthreads=0):
        L = load_host_library()
        self.pos = np.array(pos, dtype=np.float64, order="C", copy=True)
        assert self.pos.ndim == 2 and self.pos.shape[1] == 3
        n = self.pos.shape[0]
        self.perm = np.arange(n, dtype=np.int64)
        self.maxleaf = int(maxleaf)
        bdl = np.ascontiguousarray(bdl, np.float64)
        bdr = np.ascontiguousarray(bdr, np.float64)
        h = C.c_void_p()
        rc = L.p2p_build_localtree(C.byref(h), self.pos.ctypes.data_as(_dp), 3, self.perm.ctypes.data_as(_lp), n,
                                   self.maxleaf, bdl.ctypes.data_as(_dp), bdr.ctypes.data_as(_dp), int(direct_start),
                                   int(nthreads))
        if rc != 0:
            raise P2PError(rc, "p2p_build_localtree failed (capacity 2N/MAXLEAF exceeded?)")
        self._h = h
        v = _TreeView()
        L.p2p_tree_get(h, C.byref(v))
        self.npart, self.nleaf, self.nnode = v.npart, v.nleaf, v.nnode
        self.nleaf_cap, self.first_leaf, self.first_node = v.nleaf_cap, v.first_leaf, v.first_node

        def arr(p, shape, dt):
            if int(np.prod(shape)) == 0:
                return np.empty(shape, dt)
            return np.ctypeslib.as_array(p, shape=shape).astype(dt, copy=True)

        self.leaf_npart = arr(v.leaf_npart, (v.nleaf,), np.int32)
        self.leaf_ipart = arr(v.leaf_ipart, (v.nleaf,), np.int32)
        self.leaf_center = arr(v.leaf_center, (v.nleaf, 3), np.float64)
        self.leaf_width = arr(v.leaf_width, (v.nleaf, 3), np.float64)
        self.node_npart = arr(v.node_npart, (v.nnode,), np.int32)
        self.node_son = arr(v.node_son, (v.nnode, 2), np.int32)
        self.node_split = arr(v.node_split, (v.nnode,), np.float64)
        self.node_center = arr(v.node_center, (v.nnode, 3), np.float64)
        self.node_width = arr(v.node_width, (v.nnode, 3), np.float64)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load_host_library().p2p_tree_free(self._h)
                self._h = None
        except Exception:
            pass

    def walk_task_p2p(self, theta, rcut, nthreads=0):
        """(tt, ts): 0-based (target leaf, source leaf) in the reference's traversal order."""
        L = load_host_library()
        tt, ts, n = _ip(), _ip(), C.c_int64()
        rc = L.p2p_walk_task_p2p(self._h, float(theta), float(rcut), int(nthreads), C.byref(tt), C.byref(ts), C.byref(n))
        if rc != 0:
            raise P2PError(rc, "p2p_walk_task_p2p failed")
        return _take(tt, n.value, np.int32), _take(ts, n.value, np.int32)

    def walk_plan(self, theta, rcut, nchunks):
        """Chunked walk for the walk/compute pipeline (see p2p_host.h)."""
        return WalkPlan(self, theta, rcut, nchunks)

    def prepare_sendtree(self, tcenter, twidth, displace, theta, rcut):
        L = load_host_library()
        img = _Image()
        tc = np.ascontiguousarray(tcenter, np.float64)
        tw = np.ascontiguousarray(twidth, np.float64)
        dp = np.ascontiguousarray(displace, np.float64)
        rc = L.p2p_prepare_sendtree(self._h, self.pos.ctypes.data_as(_dp), 3, tc.ctypes.data_as(_dp), tw.ctypes.data_as(_dp),
                                    dp.ctypes.data_as(_dp), float(theta), float(rcut), C.byref(img))
        if rc != 0:
            raise P2PError(rc, "p2p_prepare_sendtree failed")
        nn, nb = img.nnode, img.nbody

        def arr(p, shape, dt):
            if int(np.prod(shape)) == 0:
                return np.empty(shape, dt)
            return np.ctypeslib.as_array(p, shape=shape).astype(dt, copy=True)

        out = Image(arr(img.npart, (nn,), np.int32), arr(img.son, (nn, 2), np.int32), arr(img.center, (nn, 3), np.float64),
                    arr(img.width, (nn, 3), np.float64), arr(img.body, (nb, 3), np.float64))
        L.p2p_image_free(C.byref(img))
        return out

    def walk_task_p2p_ext(self, image, theta, rcut, nthreads=0):
        """(tt 0-based local leaf, ts image node index)."""
        L = load_host_library()
        ci = image._c()
        tt, ts, n = _ip(), _ip(), C.c_int64()
        rc = L.p2p_walk_task_p2p_ext(self._h, C.byref(ci), float(theta), float(rcut), int(nthreads), C.byref(tt),
                                     C.byref(ts), C.byref(n))
        if rc != 0:
            raise P2PError(rc, "p2p_walk_task_p2p_ext failed")
        return _take(tt, n.value, np.int32), _take(ts, n.value, np.int32)


class WalkPlan:
    """Frontier of the dual-tree recursion grouped by target chunk; run(c) walks one chunk."""

    def __init__(self, tree, theta, rcut, nchunks):
        L = load_host_library()
        self._tree = tree                      # keeps the tree alive
        h = C.c_void_p()
        rc = L.p2p_walk_plan_create(tree._h, float(theta), float(rcut), int(nchunks), C.byref(h))
        if rc != 0:
            raise P2PError(rc, "p2p_walk_plan_create failed")
        self._h = h
        self.nchunks = L.p2p_walk_plan_nchunks(h)

    def rows(self, c):
        b, e = C.c_int(), C.c_int()
        load_host_library().p2p_walk_plan_rows(self._h, int(c), C.byref(b), C.byref(e))
        return b.value, e.value

    def run(self, c, nthreads=0):
        L = load_host_library()
        tt, ts, n = _ip(), _ip(), C.c_int64()
        rc = L.p2p_walk_plan_run(self._h, int(c), int(nthreads), C.byref(tt), C.byref(ts), C.byref(n))
        if rc != 0:
            raise P2PError(rc, "p2p_walk_plan_run failed")
        return _take(tt, n.value, np.int32), _take(ts, n.value, np.int32)

    def run_into(self, c, tt_buf, ts_buf, nthreads=0):
        """Walk chunk c straight into caller-owned int32 buffers (pinned memory); returns the task count,
        or -needed if the buffers are too small."""
        L = load_host_library()
        n = C.c_int64()
        rc = L.p2p_walk_plan_run_into(self._h, int(c), int(nthreads), tt_buf.ctypes.data_as(_ip), ts_buf.ctypes.data_as(_ip),
                                      min(len(tt_buf), len(ts_buf)), C.byref(n))
        if rc == -3:
            return -n.value
        if rc != 0:
            raise P2PError(rc, "p2p_walk_plan_run_into failed")
        return n.value

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load_host_library().p2p_walk_plan_free(self._h)
                self._h = None
        except Exception:
            pass


def domain_setup(nproc, box):
    L = load_host_library()
    n = 2 * nproc - 1
    split, center, width, direct = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n, np.int32)
    rc = L.p2p_domain_setup(int(nproc), float(box), split.ctypes.data_as(_dp), center.ctypes.data_as(_dp),
                            width.ctypes.data_as(_dp), direct.ctypes.data_as(_ip))
    if rc != 0:
        raise P2PError(rc, "p2p_domain_setup failed")
    return split, center, width, direct


def domain_of_rank(nproc, rank):
    return int(load_host_library().p2p_domain_of_rank(int(nproc), int(rank)))


def domain_route(nproc, split, pos, payload):
    """In-place routing of (pos, payload) by the rank kd-tree; returns sendcount[nproc]."""
    L = load_host_library()
    assert pos.dtype == np.float64 and pos.flags.c_contiguous and payload.dtype == np.int64
    send = np.zeros(nproc, np.int32)
    rc = L.p2p_domain_route(int(nproc), np.ascontiguousarray(split).ctypes.data_as(_dp), pos.ctypes.data_as(_dp),
                            pos.shape[1], payload.ctypes.data_as(_lp), pos.shape[0], send.ctypes.data_as(_ip))
    if rc != 0:
        raise P2PError(rc, "p2p_domain_route failed")
    return send


def domain_relax(nproc, box, split, work):
    """Splits of the next step from this step's per-rank task counts (the reference's load-balance feedback)."""
    out = np.array(split, np.float64, copy=True)
    w = np.ascontiguousarray(work, np.float64)
    rc = load_host_library().p2p_domain_relax(int(nproc), float(box), out.ctypes.data_as(_dp), w.ctypes.data_as(_dp))
    if rc != 0:
        raise P2PError(rc, "p2p_domain_relax failed")
    return out


def domain_boxes(nproc, box, split):
    """(center, width, direct) of the rank-tree nodes for given splits (e.g. after domain_relax)."""
    n = 2 * nproc - 1
    center, width, direct = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n, np.int32)
    sp = np.ascontiguousarray(split, np.float64)
    rc = load_host_library().p2p_domain_boxes(int(nproc), float(box), sp.ctypes.data_as(_dp), center.ctypes.data_as(_dp),
                                              width.ctypes.data_as(_dp), direct.ctypes.data_as(_ip))
    if rc != 0:
        raise P2PError(rc, "p2p_domain_boxes failed")
    return center, width, direct


def max_threads():
    return int(load_host_library().p2p_host_max_threads())
