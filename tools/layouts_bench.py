"""the same (target leaf, source leaf) list through the three list layouts --
Indexing ABI (index pairs into a leaf-chunked padded position array, per-TASK result slots; 1_Indexing/src/fmm.c:842-911),
Redundant ABI (private copies of targets and sources per task; 2_Redundant/src/fmm.c:790-881) and the native CSR.
Times the three blocking ABI calls the reference makes (copy, launch, read) and the caller's host-side reduction of the
per-task slots; packing of the reference layouts is numpy here and reported separately.
usage: python tools/layouts_bench.py [nside=64] [maxleaf=16]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200  # noqa: E402
from p2p_b200 import host, synth  # noqa: E402

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 64
maxleaf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pos, box = synth.zeldovich_like(nside)
rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
T = host.LocalTree(pos, maxleaf, [0.0] * 3, [box] * 3, 0)
tt, ts = T.walk_task_p2p(0.4, rcut)
nt = len(tt)
nT, nS = T.leaf_npart[tt].astype(np.int64), T.leaf_npart[ts].astype(np.int64)
npairs = int((nT * nS).sum())
maxp = int(T.leaf_npart.max())
out = {"workload": f"{nside}^3 Zeldovich-like, MAXLEAF {maxleaf}, local list", "particles": int(T.npart), "leaves": int(T.nleaf),
       "tasks": nt, "pairs": npairs, "max_leaf": maxp}


def tm(f):
    t0 = time.perf_counter()
    r = f()
    return r, time.perf_counter() - t0


def reduce_slots(result):
    """the caller's update loop (1_Indexing/src/fmm.c:895-908), vectorised"""
    acc = np.zeros((T.npart, 3))
    for i in range(maxp):
        m = nT > i
        np.add.at(acc, T.leaf_ipart[tt[m]] + i, result[m, i])
    return acc


# ---------------------------------------------------------------- native CSR
ctx = p2p_b200.P2PContext(0)
ctx.set_physics(1.0, eps, rs)
ctx.set_box([0.0, 0.0, 0.0], box)
for _ in range(3):
    acc_csr, dt = tm(lambda: ctx.step_host(T.pos, T.leaf_npart, T.leaf_ipart, tt, ts))
ms_k, ms_csr = ctx.last_timings()
out["csr"] = {"total_s": dt, "kernel_ms": ms_k, "pack_ms": ms_csr, "pair_per_s_total": npairs / dt, "pair_per_s_kernel": npairs / (ms_k * 1e-3),
              "h2d_bytes": int(T.pos.nbytes + 8 * T.nleaf + 8 * nt), "d2h_bytes": int(acc_csr.nbytes)}
ctx.close()

# ---------------------------------------------------------------- Indexing layout
L = C.CDLL(os.path.join(p2p_b200.LIB_DIR, "libphotoNs_CUDA_indexing.so"))
L.copyMemGPU.argtypes = [_dp, _ip, _ip, C.c_int, C.c_int]
L.LaunchKernelP2PIndexing.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double, C.c_int]
L.readResultsGPU.argtypes = [_dp, C.c_int, C.c_int, C.c_int]
L.p2pSetSplitRadius.argtypes = [C.c_double]
t0 = time.perf_counter()
particle_data = np.full((T.nleaf, maxp, 3), np.nan)
leaf_data = np.stack([T.leaf_npart, T.leaf_ipart], axis=1).astype(np.int32).copy()
for k in range(maxp):
    m = T.leaf_npart > k
    particle_data[m, k] = T.pos[T.leaf_ipart[m] + k]
inter = np.stack([tt, ts], axis=1).astype(np.int32).copy()
t_pack = time.perf_counter() - t0
L.p2pSetSplitRadius(rs)
L.initGPU(0)
assert L.allocMemGPU(T.nleaf, maxp, 1000, nt, 0) == 0
result = np.empty((nt, maxp, 3))
for _ in range(2):
    _, t_copy = tm(lambda: L.copyMemGPU(particle_data.ctypes.data_as(_dp), leaf_data.ctypes.data_as(_ip), inter.ctypes.data_as(_ip), nt, 0))
    _, t_launch = tm(lambda: L.LaunchKernelP2PIndexing(nt, maxp * 3, 2, maxp * 3, eps, 1.0, 0))
    _, t_read = tm(lambda: L.readResultsGPU(result.ctypes.data_as(_dp), nt, maxp, 0))
acc_idx, t_red = tm(lambda: reduce_slots(result))
tot = t_copy + t_launch + t_read
out["indexing_abi"] = {"copyMemGPU_s": t_copy, "launch_sync_s": t_launch, "readResultsGPU_s": t_read, "abi_total_s": tot,
                       "host_reduction_s_numpy": t_red, "host_pack_s_numpy": t_pack, "pair_per_s_abi": npairs / tot,
                       "pair_per_s_launch": npairs / t_launch, "h2d_bytes": int(particle_data.nbytes + leaf_data.nbytes + inter.nbytes),
                       "d2h_bytes": int(result.nbytes), "max_abs_diff_vs_csr": float(np.abs(acc_idx - acc_csr).max() / np.abs(acc_csr).max())}
del result, particle_data

# ---------------------------------------------------------------- Redundant layout (private copies per task)
R = C.CDLL(os.path.join(p2p_b200.LIB_DIR, "libphotoNs_CUDA_redundant.so"))
R.allocAndCopySelfInteractionsGPU.argtypes = [_dp, _ip, C.c_int, C.c_int, C.c_int, C.c_int]
R.LaunchKernelP2PSelfInteractions.argtypes = [C.c_int] * 4 + [C.c_double, C.c_double]
R.readResultsGPUSelfInteractions.argtypes = [_dp, C.c_int, C.c_int]
R.p2pSetSplitRadius.argtypes = [C.c_double]
chunk, rchunk = maxp * 2 * 3, maxp * 3
t0 = time.perf_counter()
part_data = np.zeros((nt, chunk))
for k in range(maxp):
    mt = nT > k
    part_data[mt, 3 * k:3 * k + 3] = T.pos[T.leaf_ipart[tt[mt]] + k]
for k in range(maxp):
    idx = np.nonzero(nS > k)[0]
    col = 3 * (nT[idx] + k)
    src = T.pos[T.leaf_ipart[ts[idx]] + k]
    for c in range(3):
        part_data[idx, col + c] = src[:, c]
part_idx = np.stack([nT, nS, tt + T.first_leaf], axis=1).astype(np.int32).copy()
t_pack = time.perf_counter() - t0
R.p2pSetSplitRadius(rs)
R.initGPU(0)
result = np.zeros((nt, maxp, 3))
for _ in range(2):
    _, t_copy = tm(lambda: R.allocAndCopySelfInteractionsGPU(part_data.ctypes.data_as(_dp), part_idx.ctypes.data_as(_ip), chunk, 3, rchunk, nt))
    _, t_launch = tm(lambda: R.LaunchKernelP2PSelfInteractions(nt, chunk, 3, rchunk, eps, 1.0))
    _, t_read = tm(lambda: R.readResultsGPUSelfInteractions(result.ctypes.data_as(_dp), rchunk, nt))
acc_red, t_red = tm(lambda: reduce_slots(result))
tot = t_copy + t_launch + t_read
kernel_bytes = float(((nT + nS) * 24 + nT * 24).sum())
out["redundant_abi"] = {"allocAndCopy_s": t_copy, "launch_sync_s": t_launch, "readResults_s": t_read, "abi_total_s": tot,
                        "host_reduction_s_numpy": t_red, "host_pack_s_numpy": t_pack, "pair_per_s_abi": npairs / tot,
                        "pair_per_s_launch": npairs / t_launch, "kernel_algorithmic_bytes": kernel_bytes,
                        "kernel_gbs": kernel_bytes / t_launch / 1e9, "h2d_bytes": int(part_data.nbytes + part_idx.nbytes),
                        "d2h_bytes": int(result.nbytes), "max_abs_diff_vs_csr": float(np.abs(acc_red - acc_csr).max() / np.abs(acc_csr).max())}
print(json.dumps(out))
