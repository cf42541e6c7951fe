"""ctypes front-end of oracle/libp2p_oracle.so (the fp64 CPU restatement of the reference's P2P path).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libp2p_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.isfile(LIB_PATH):
        subprocess.run(["make", "-C", HERE, "libp2p_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", HERE, "ref"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_walk_p2p.restype = C.c_int64
        _lib.oracle_walk_p2p_ext.restype = C.c_int64
        _lib.oracle_p2p_tasks.restype = C.c_int64
        _lib.oracle_p2p_absterms.restype = C.c_int64
        _lib.oracle_fingerprint.restype = C.c_uint64
        _lib.oracle_midfield.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _l(a):
    return a.ctypes.data_as(_lp)


class Tree:
    """Local kd-tree in the reference's layout (I/src/fmm.c:176-263)."""

    def __init__(self, pos, maxleaf, bdl, bdr, direct_start=0):
        pos = np.array(pos, dtype=np.float64, order="C", copy=True)
        n = pos.shape[0]
        self.npart, self.maxleaf = n, int(maxleaf)
        nl, nn = C.c_int(), C.c_int()
        lib().oracle_tree_caps(n, self.maxleaf, C.byref(nl), C.byref(nn))
        self.nleaf_cap, self.nnode_cap = nl.value, nn.value
        self.first_leaf, self.first_node = n, n + self.nleaf_cap
        self.pos = pos
        self.perm = np.arange(n, dtype=np.int64)
        self.leaf_npart = np.zeros(self.nleaf_cap, np.int32)
        self.leaf_ipart = np.zeros(self.nleaf_cap, np.int32)
        self.leaf_center = np.zeros((self.nleaf_cap, 3))
        self.leaf_width = np.zeros((self.nleaf_cap, 3))
        self.node_npart = np.zeros(self.nnode_cap, np.int32)
        self.node_son = np.zeros((self.nnode_cap, 2), np.int32)
        self.node_split = np.zeros(self.nnode_cap)
        self.node_center = np.zeros((self.nnode_cap, 3))
        self.node_width = np.zeros((self.nnode_cap, 3))
        bdl = np.asarray(bdl, np.float64)
        bdr = np.asarray(bdr, np.float64)
        rc = lib().oracle_build_localtree(
            n, self.maxleaf, int(direct_start), _d(bdl), _d(bdr), _d(self.pos), _l(self.perm),
            _i(self.leaf_npart), _i(self.leaf_ipart), _d(self.leaf_center), _d(self.leaf_width),
            _i(self.node_npart), _i(self.node_son), _d(self.node_split), _d(self.node_center),
            _d(self.node_width), C.byref(nl), C.byref(nn))
        if rc != 0:
            raise RuntimeError("oracle_build_localtree: capacity exceeded")
        self.nleaf, self.nnode = nl.value, nn.value

    def walk_p2p(self, theta, rcut, cap=None):
        """(tt, ts) 0-based leaf indices in the reference's traversal order (I/src/fmm.c:402-534)."""
        cap = int(cap if cap is not None else max(1024, self.nleaf * 400))
        while True:
            tt = np.empty(cap, np.int32)
            ts = np.empty(cap, np.int32)
            n = lib().oracle_walk_p2p(self.npart, self.nleaf_cap, self.nleaf, self.nnode, _d(self.leaf_center),
                                      _d(self.leaf_width), _i(self.node_son), _d(self.node_center),
                                      _d(self.node_width), C.c_double(theta), C.c_double(rcut), _i(tt), _i(ts),
                                      C.c_int64(cap))
            if n <= cap:
                return tt[:n] - self.first_leaf, ts[:n] - self.first_leaf
            cap = int(n)

    def prune(self, tcenter, twidth, displace, theta, rcut):
        """Pruned send tree + displaced leaf bodies for one target domain box (I/src/remotes.c:337-446)."""
        cap_node = self.nleaf + self.nnode + 2
        cap_body = self.npart + 1
        r_npart = np.zeros(cap_node, np.int32)
        r_son = np.zeros((cap_node, 2), np.int32)
        r_center = np.zeros((cap_node, 3))
        r_width = np.zeros((cap_node, 3))
        r_body = np.zeros((cap_body, 3))
        nn, nb = C.c_int(), C.c_int()
        tc = np.asarray(tcenter, np.float64)
        tw = np.asarray(twidth, np.float64)
        dp = np.asarray(displace, np.float64)
        rc = lib().oracle_prune_sendtree(
            self.npart, self.nleaf_cap, _d(self.pos), _i(self.leaf_npart), _i(self.leaf_ipart), _d(self.leaf_center),
            _d(self.leaf_width), _i(self.node_npart), _i(self.node_son), _d(self.node_center), _d(self.node_width),
            _d(tc), _d(tw), _d(dp), C.c_double(theta), C.c_double(rcut), cap_node, cap_body, _i(r_npart), _i(r_son),
            _d(r_center), _d(r_width), _d(r_body), C.byref(nn), C.byref(nb))
        if rc != 0:
            raise RuntimeError("oracle_prune_sendtree overflow")
        n, b = nn.value, nb.value
        return dict(npart=r_npart[:n].copy(), son=r_son[:n].copy(), center=r_center[:n].copy(),
                    width=r_width[:n].copy(), body=r_body[:b].copy())

    def walk_p2p_ext(self, remote, theta, rcut, cap=None):
        """(tt 0-based local leaf, ts remote node index) (I/src/remotes.c:141-317)."""
        nrn = len(remote["npart"])
        if nrn == 0:
            return np.empty(0, np.int32), np.empty(0, np.int32)
        cap = int(cap if cap is not None else max(1024, self.nleaf * 200))
        r_npart = np.ascontiguousarray(remote["npart"], np.int32)
        r_son = np.ascontiguousarray(remote["son"], np.int32)
        r_center = np.ascontiguousarray(remote["center"], np.float64)
        r_width = np.ascontiguousarray(remote["width"], np.float64)
        while True:
            tt = np.empty(cap, np.int32)
            ts = np.empty(cap, np.int32)
            n = lib().oracle_walk_p2p_ext(self.npart, self.nleaf_cap, self.maxleaf, _d(self.leaf_center),
                                          _d(self.leaf_width), _i(self.node_son), _d(self.node_center),
                                          _d(self.node_width), nrn, _i(r_npart), _i(r_son), _d(r_center),
                                          _d(r_width), C.c_double(theta), C.c_double(rcut), _i(tt), _i(ts),
                                          C.c_int64(cap))
            if n <= cap:
                return tt[:n] - self.first_leaf, ts[:n].copy()
            cap = int(n)


def midfield(T, theta, rcut, rs, mass, box=0.0, literal_d6=False):
    """Mid-field of a single rank's tree (oracle.Tree): P2M / M2M / M2L / L2L / L2P as the reference runs them
    (1_Indexing/src/operator.c).  Returns dict(leaf_M, node_M, leaf_L, acc [tree order], nm2l_local, nm2l_total)."""
    nl, nn = T.nleaf, T.nnode
    leaf_M, node_M, leaf_L = np.zeros((nl, 20)), np.zeros((nn, 20)), np.zeros((nl, 20))
    acc = np.zeros((T.npart, 3))
    a, b = C.c_int64(), C.c_int64()
    rc = lib().oracle_midfield(T.npart, T.nleaf_cap, nl, nn, T.maxleaf, _d(T.pos), _i(T.leaf_npart), _i(T.leaf_ipart), _d(T.leaf_center),
                               _d(T.leaf_width), _i(T.node_npart), _i(T.node_son), _d(T.node_center), _d(T.node_width),
                               C.c_double(theta), C.c_double(rcut), C.c_double(rs), C.c_double(mass), C.c_double(box),
                               1 if literal_d6 else 0, _d(leaf_M), _d(node_M), _d(leaf_L), _d(acc), C.byref(a), C.byref(b))
    if rc != 0:
        raise RuntimeError("oracle_midfield failed")
    return dict(leaf_M=leaf_M, node_M=node_M, leaf_L=leaf_L, acc=acc, nm2l_local=a.value, nm2l_total=b.value)


def p2p(tpos, t_npart, t_ipart, spos, s_count, s_start, tt, ts, mass, eps, rs, acc=None, nthreads=0, absterms=False):
    """fp64 P2P over a task list; returns (acc[N,3], npairs).  rs <= 0 selects the plain kernel."""
    tpos = np.ascontiguousarray(tpos, np.float64)
    spos = np.ascontiguousarray(spos, np.float64)
    t_npart = np.ascontiguousarray(t_npart, np.int32)
    t_ipart = np.ascontiguousarray(t_ipart, np.int32)
    s_count = np.ascontiguousarray(s_count, np.int32)
    s_start = np.ascontiguousarray(s_start, np.int32)
    tt = np.ascontiguousarray(tt, np.int32)
    ts = np.ascontiguousarray(ts, np.int32)
    if acc is None:
        acc = np.zeros((tpos.shape[0], 3))
    fn = lib().oracle_p2p_absterms if absterms else lib().oracle_p2p_tasks
    npairs = fn(_d(tpos), _i(t_npart), _i(t_ipart), len(t_npart), _d(spos), _i(s_count), _i(s_start), _i(tt), _i(ts),
                C.c_int64(len(tt)), C.c_double(mass), C.c_double(eps), C.c_double(rs), _d(acc), int(nthreads))
    return acc, int(npairs)


def fingerprint_sorted(tt, ts):
    """SURVEY.md section 8c fingerprint of the list sorted by (t, s)."""
    order = np.lexsort((ts, tt))
    stream = np.empty(2 * len(tt), np.int32)
    stream[0::2] = np.asarray(tt)[order]
    stream[1::2] = np.asarray(ts)[order]
    return int(lib().oracle_fingerprint(stream.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(len(stream))))


def derived_params(box, nside, npart_total):
    """I/src/initial.c:324-346: splitRadius, cutoffRadius, SoftenScale."""
    rs = 1.25 * (box / float(nside))
    eps = 0.03 * box / pow(float(npart_total), 0.3333333)
    return rs, 4.5 * rs, eps


def domain_setup(nproc, box):
    n = 2 * nproc - 1
    split = np.zeros(n)
    center = np.zeros((n, 3))
    width = np.zeros((n, 3))
    direct = np.zeros(n, np.int32)
    lib().oracle_domain_setup(int(nproc), C.c_double(box), _d(split), _d(center), _d(width), _i(direct))
    return split, center, width, direct


def domain_of_rank(nproc, rank):
    return int(lib().oracle_domain_of_rank(int(nproc), int(rank)))


def domain_partition(nproc, split, pos, payload):
    """In-place routing partition (I/src/domains.c:272-296); returns sendcount[nproc]."""
    send = np.zeros(nproc, np.int32)
    lib().oracle_domain_partition(int(nproc), _d(split), _d(pos), _l(payload), pos.shape[0], _i(send), None)
    return send


def domain_relax(nproc, box, split, work):
    """New splits after the reference's load-balance feedback; work[r] = tasks of rank r (idxP2P + idxM2L)."""
    work = np.asarray(work, np.float64)
    frac = np.ascontiguousarray(work * nproc / (work.sum() + 0.0001))        # I/src/photoNs.c:303
    out = np.array(split, np.float64, copy=True)
    lib().oracle_domain_relax(int(nproc), C.c_double(box), _d(out), _d(frac))
    return out


def max_threads():
    return int(lib().oracle_max_threads())
