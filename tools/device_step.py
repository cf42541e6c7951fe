"""Times the device-resident short-range step (tree build, dual-tree walk, packing, forces) on one GPU.
usage: python tools/device_step.py [nside] [maxleaf] [reps] [--clustered] [--midfield] [--theta=X] [--resident]
--resident: also time device-resident steps (forces + kick + drift with the particles staying in HBM)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200  # noqa: E402
from p2p_b200 import host, synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nside = int(args[0]) if len(args) > 0 else 128
maxleaf = int(args[1]) if len(args) > 1 else 32
reps = int(args[2]) if len(args) > 2 else 3
pos, box = (synth.clustered(nside) if "--clustered" in sys.argv else synth.zeldovich_like(nside))
rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
ctx = p2p_b200.P2PContext(0)
ctx.set_physics(1.0, eps, rs)
ctx.set_box([0.0, 0.0, 0.0], box)
bdl, bdr = np.zeros(3), np.full(3, box)
import torch  # noqa: E402  (pinned host buffers)
ppos = torch.from_numpy(pos).pin_memory().numpy()
acc = torch.empty((pos.shape[0], 3), dtype=torch.float64).pin_memory().numpy()
theta = max([float(a.split("=")[1]) for a in sys.argv if a.startswith("--theta=")] + [0.0]) or 0.4
mid = "--midfield" in sys.argv
ctx.midfield_enable(mid)
for r in range(reps):
    t0 = time.perf_counter()
    ctx.tree_build(ppos, maxleaf, bdl, bdr, 0)
    t1 = time.perf_counter()
    ctx.clear_tasks()
    ctx.tree_walk(theta, rcut, box, 0.5 * (bdr + bdl), bdr - bdl)
    t2 = time.perf_counter()
    ctx.build_csr()
    ctx.compute()
    nm2l = ctx.midfield_compute() if mid else 0
    ctx.download_acc_original(acc)
    t3 = time.perf_counter()
    info = ctx.tree_info()
    ms_k, ms_csr = ctx.last_timings()
    nt, npairs = ctx.counts()
    print(f"rep {r}: total {1e3 * (t3 - t0):.1f} ms | build {1e3 * (t1 - t0):.1f} (device {info['ms_build']:.1f}) "
          f"walk {1e3 * (t2 - t1):.1f} (device {info['ms_walk']:.1f}, {info['walk_items']} items) "
          f"csr {ms_csr:.1f} force {ms_k:.1f} rest {1e3 * (t3 - t2) - ms_csr - ms_k:.1f} | "
          + (f"midfield {ctx.midfield_download()['ms']:.2f} ms ({nm2l} M2L tasks) | " if mid else "") +
          f"{info['nleaf']} leaves {info['nlevel']} levels {nt} tasks {npairs} pairs dup {ctx.csr_duplicates()}", flush=True)

if "--resident" in sys.argv:
    ctx.midfield_enable(False)
    ctx.resident_load(ppos)
    cell = box / nside
    for r in range(reps + 1):
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.resident_forces(maxleaf, bdl, bdr, theta, rcut, box)
        info = ctx.tree_info()                  # (host-side numbers only)
        ctx.resident_kick(1e-3 * cell)          # small steps: the box stays quasi-uniform
        ctx.resident_drift(1e-3, box)
        ctx.synchronize()
        t1 = time.perf_counter()
        ms_k, ms_csr = ctx.last_timings()
        print(f"resident step {r}: {1e3 * (t1 - t0):.1f} ms | build {info['ms_build']:.1f} walk {info['ms_walk']:.1f} csr {ms_csr:.1f} force {ms_k:.1f} | "
              f"{ctx.accumulated_counts()[1]} pairs", flush=True)
