"""Drop-in test of the compat shims: the reference's own symbol names, signatures and data layouts
(1_Indexing/inc/photoNs_CUDA.cuh:24-33, 2_Redundant/inc/photoNs_CUDA.cuh:28-51), driven exactly the
way task_compute_p2p packs them (1_Indexing/src/fmm.c:842-911, 2_Redundant/src/fmm.c:790-881)."""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_MASS, DEMO_NSIDE, THETA

import oracle
import p2p_b200

pytestmark = pytest.mark.gpu
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _demo_lists(demo_pos, maxleaf):
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    T = oracle.Tree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = T.walk_p2p(THETA, rcut)
    return T, tt, ts, rs, eps


@pytest.mark.parametrize("truncated", [False, True])
def test_indexing_abi(demo_pos, truncated):
    L = C.CDLL(os.path.join(p2p_b200.LIB_DIR, "libphotoNs_CUDA_indexing.so"))
    L.copyMemGPU.argtypes = [_dp, _ip, _ip, C.c_int, C.c_int]
    L.LaunchKernelP2PIndexing.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double, C.c_int]
    L.readResultsGPU.argtypes = [_dp, C.c_int, C.c_int, C.c_int]
    L.p2pSetSplitRadius.argtypes = [C.c_double]
    T, tt, ts, rs, eps = _demo_lists(demo_pos, 16)
    maxp = int(T.leaf_npart[:T.nleaf].max())
    # pack exactly like task_compute_p2p (1_Indexing/src/fmm.c:851-877)
    # padding slots hold stale memory in the reference (never read by its kernel): NaN in one run, huge finite values in the other
    particle_data = np.full((T.nleaf_cap, maxp, 3), 1e30 if truncated else np.nan)
    leaf_data = np.zeros((T.nleaf_cap, 2), np.int32)
    for l in range(T.nleaf):
        n, ip = T.leaf_npart[l], T.leaf_ipart[l]
        leaf_data[l] = (n, ip)
        particle_data[l, :n] = T.pos[ip:ip + n]
    inter = np.stack([tt, ts], axis=1).astype(np.int32).copy()
    nt = len(tt)
    L.p2pSetSplitRadius(rs if truncated else 0.0)
    L.initGPU(0)
    assert L.allocMemGPU(T.nleaf_cap, maxp, 1000, T.nleaf_cap * 1000, 0) == 0
    assert L.copyMemGPU(particle_data.ctypes.data_as(_dp), leaf_data.ctypes.data_as(_ip), inter.ctypes.data_as(_ip), nt, 0) == 0
    assert L.LaunchKernelP2PIndexing(nt, maxp * 3, 2, maxp * 3, eps, DEMO_MASS, 0) == 0
    result = np.empty((nt, maxp, 3))
    L.readResultsGPU(result.ctypes.data_as(_dp), nt, maxp, 0)
    # the caller's update loop (1_Indexing/src/fmm.c:895-908)
    acc = np.zeros((T.npart, 3))
    n_t = T.leaf_npart[tt]
    for i in range(maxp):
        m = n_t > i
        np.add.at(acc, T.leaf_ipart[tt[m]] + i, result[m, i])
    ref, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, DEMO_MASS, eps,
                        rs if truncated else 0.0)
    nr = np.linalg.norm(ref, axis=1)
    err = (np.linalg.norm(acc - ref, axis=1) / np.maximum(nr, nr.mean())).max()
    assert err < 1e-5, err
    # remote-style call with ids outside the leaf table (SURVEY defect D3) is refused, not executed
    bad = inter.copy()
    bad[0, 1] = -5
    assert L.copyMemGPU(particle_data.ctypes.data_as(_dp), leaf_data.ctypes.data_as(_ip), bad.ctypes.data_as(_ip), nt, 0) == -3


def test_redundant_abi(demo_pos):
    L = C.CDLL(os.path.join(p2p_b200.LIB_DIR, "libphotoNs_CUDA_redundant.so"))
    L.allocAndCopySelfInteractionsGPU.argtypes = [_dp, _ip, C.c_int, C.c_int, C.c_int, C.c_int]
    L.LaunchKernelP2PSelfInteractions.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double]
    L.readResultsGPUSelfInteractions.argtypes = [_dp, C.c_int, C.c_int]
    L.p2pSetSplitRadius.argtypes = [C.c_double]
    T, tt, ts, rs, eps = _demo_lists(demo_pos, 8)
    sel = tt < 800                                # complete rows of the first 800 target leaves (~200k tasks)
    tt, ts = tt[sel], ts[sel]
    nt = len(tt)
    maxp = 8
    chunk, rchunk = maxp * 2 * 3, maxp * 3
    part_data = np.zeros((nt, chunk))
    part_idx = np.zeros((nt, 3), np.int32)
    for k in range(maxp):                          # private copies: targets then sources (intended layout, D9 fixed)
        mt = T.leaf_npart[tt] > k
        part_data[mt, 3 * k:3 * k + 3] = T.pos[T.leaf_ipart[tt[mt]] + k]
    nT = T.leaf_npart[tt]
    for k in range(maxp):
        ms = T.leaf_npart[ts] > k
        idx = np.nonzero(ms)[0]
        col = 3 * (nT[idx] + k)
        src = T.pos[T.leaf_ipart[ts[idx]] + k]
        for c in range(3):
            part_data[idx, col + c] = src[:, c]
    part_idx[:, 0], part_idx[:, 1], part_idx[:, 2] = nT, T.leaf_npart[ts], tt + T.first_leaf
    L.p2pSetSplitRadius(rs)
    L.initGPU(0)
    assert L.allocAndCopySelfInteractionsGPU(part_data.ctypes.data_as(_dp), part_idx.ctypes.data_as(_ip), chunk, 3, rchunk, nt) == 0
    L.LaunchKernelP2PSelfInteractions(nt, chunk, 3, rchunk, eps, DEMO_MASS)
    result = np.zeros((nt, maxp, 3))
    L.readResultsGPUSelfInteractions(result.ctypes.data_as(_dp), rchunk, nt)
    acc = np.zeros((T.npart, 3))
    for i in range(maxp):
        m = nT > i
        np.add.at(acc, T.leaf_ipart[tt[m]] + i, result[m, i])
    ref, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, DEMO_MASS, eps, rs)
    absr, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, DEMO_MASS, eps, rs, absterms=True)
    d = np.linalg.norm(acc - ref, axis=1)
    assert (d / np.maximum(np.linalg.norm(absr, axis=1), 1e-300)).max() < 1e-5

    # remote (DualNaive) layout: variable-length private copies + 5-int descriptors (2_Redundant/src/remotes.c:55-98)
    L.copyMemGPU.argtypes = [C.POINTER(_dp), C.POINTER(_ip), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.LaunchKernelP2PDualNaive.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    L.readResultsGPU.argtypes = [C.POINTER(_dp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    nt2 = 50000
    tt2, ts2 = tt[:nt2], ts[:nt2]
    nT2, nS2 = T.leaf_npart[tt2], T.leaf_npart[ts2]
    sizes = 3 * (nT2 + nS2)
    start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    pos_data = np.zeros(int(sizes.sum()))
    for n in range(nt2):
        a = T.pos[T.leaf_ipart[tt2[n]]: T.leaf_ipart[tt2[n]] + nT2[n]].ravel()
        b = T.pos[T.leaf_ipart[ts2[n]]: T.leaf_ipart[ts2[n]] + nS2[n]].ravel()
        pos_data[start[n]: start[n] + len(a)] = a
        pos_data[start[n] + len(a): start[n] + len(a) + len(b)] = b
    index = np.stack([start, tt2 + T.first_leaf, nT2, nS2, np.arange(nt2, dtype=np.int32) * maxp * 3], axis=1).astype(np.int32).copy()
    pp = (_dp * 1)(pos_data.ctypes.data_as(_dp))
    ii = (_ip * 1)(index.ctypes.data_as(_ip))
    assert L.allocMemGPU(1, maxp, nt2, 0) == 0
    assert L.copyMemGPU(pp, ii, 1, 0, maxp, nt2, len(pos_data), 0) == 0
    assert L.LaunchKernelP2PDualNaive(1, 0, nt2, eps, DEMO_MASS, 0) == 0
    res2 = np.zeros(nt2 * maxp * 3)
    rr = (_dp * 1)(res2.ctypes.data_as(_dp))
    L.readResultsGPU(rr, 0, 1, maxp, nt2, nt2 * maxp * 3, 0)
    assert np.allclose(res2.reshape(nt2, maxp, 3), result[:nt2], rtol=0, atol=1e-12 * np.abs(result).max())


@pytest.mark.parametrize("truncated", [False, True])
def test_reference_host_code_linked_against_our_library(demo_pos, tmp_path, truncated):
    """THE drop-in test: oracle/_ref/ref_dropin is the reference's unmodified fmm.c / remotes.c / toptree.c /
    domains.c / operator.c linked against lib/libphotoNs_CUDA_indexing.so (oracle/Makefile).  Its own
    fmm_prepare + fmm_task build the tree, walk it, pack the Indexing buffers, drive our C-ABI and reduce
    the result slots into part[].acc (1_Indexing/src/fmm.c:842-911)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "ref_dropin")
    if not os.path.isfile(exe):
        pytest.skip("oracle/_ref/ref_dropin not built (needs /root/reference at build time)")
    pf, of = tmp_path / "pos.f64", tmp_path / "acc.f64"
    np.ascontiguousarray(demo_pos).tofile(pf)
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    env = dict(os.environ)
    env.pop("P2P_B200_RS", None)
    if truncated:
        env["P2P_B200_RS"] = repr(rs)
    subprocess.run([exe, str(pf), str(len(demo_pos)), repr(DEMO_BOX), "16", str(DEMO_NSIDE), repr(THETA), repr(DEMO_MASS), str(of)],
                   check=True, env=env, timeout=600, stdout=subprocess.DEVNULL)
    acc = np.fromfile(of).reshape(-1, 3)
    T = oracle.Tree(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = T.walk_p2p(THETA, rcut)
    ref_t, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, DEMO_MASS, eps,
                          rs if truncated else 0.0)
    ref = np.empty_like(ref_t)
    ref[T.perm] = ref_t
    nr = np.linalg.norm(ref, axis=1)
    err = (np.linalg.norm(acc - ref, axis=1) / np.maximum(nr, nr.mean())).max()
    assert err < 1e-5, err
