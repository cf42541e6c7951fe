"""Times the device-resident short-range step (tree build, dual-tree walk, packing, forces) on one GPU.
usage: python tools/device_step.py [nside] [maxleaf] [reps] [--clustered] [--midfield] [--theta=X] [--resident] [--chunk-tasks=N]
--resident: the particles are generated slab by slab straight into HBM and stay there (every step: tree build from the
order the previous step left + walk + packing + forces); no host copy of the whole box is ever made (1024^3 = 26 GB)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200  # noqa: E402
from p2p_b200 import host, synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opt = {a.split("=")[0]: (a.split("=")[1] if "=" in a else "1") for a in sys.argv[1:] if a.startswith("--")}
nside = int(args[0]) if len(args) > 0 else 128
maxleaf = int(args[1]) if len(args) > 1 else 32
reps = int(args[2]) if len(args) > 2 else 3
clustered = "--clustered" in opt
npart = nside ** 3
box = synth.box_for(nside)
rs, rcut, eps = host.derived_params(box, nside, npart)
ctx = p2p_b200.P2PContext(0)
ctx.set_physics(1.0, eps, rs)
ctx.set_box([0.0, 0.0, 0.0], box)
if "--chunk-tasks" in opt:
    ctx.set_chunk_tasks(int(float(opt["--chunk-tasks"])))
bdl, bdr = np.zeros(3), np.full(3, box)
theta = float(opt.get("--theta", 0.4))
mid = "--midfield" in opt
ctx.midfield_enable(mid)
gen = synth.clustered_slab if clustered else synth.zeldovich_slab

if "--resident" not in opt:
    import torch  # noqa: E402  (pinned host buffers)
    pos = gen(nside, 0, npart)[0]
    ppos = torch.from_numpy(pos).pin_memory().numpy()
    acc = torch.empty((npart, 3), dtype=torch.float64).pin_memory().numpy()
    for r in range(reps):
        t0 = time.perf_counter()
        ctx.tree_build(ppos, maxleaf, bdl, bdr, 0)
        t1 = time.perf_counter()
        ctx.forces_local(theta, rcut, box, 0.5 * (bdr + bdl), bdr - bdl)
        nm2l = ctx.midfield_compute() if mid else 0
        ctx.download_acc_original(acc)
        t3 = time.perf_counter()
        info, st = ctx.tree_info(), ctx.step_timings()
        nt, npairs = ctx.accumulated_counts()
        print(f"rep {r}: total {1e3 * (t3 - t0):.1f} ms | build {1e3 * (t1 - t0):.1f} (device {st['build_ms']:.1f}) walk {st['walk_ms']:.1f} "
              f"csr {st['csr_ms']:.1f} force {st['force_ms']:.1f} chunks {st['chunks']} rest {1e3 * (t3 - t1) - st['walk_ms'] - st['csr_ms'] - st['force_ms']:.1f} | "
              + (f"midfield {ctx.midfield_download()['ms']:.2f} ms ({nm2l} M2L tasks) | " if mid else "") +
              f"{info['nleaf']} leaves {info['nlevel']} levels {nt} tasks {npairs} pairs ({npairs / st['force_ms'] / 1e9:.4f} Tpair/s in the force kernel)",
              flush=True)
else:
    t0 = time.perf_counter()
    slab = 1 << 24
    for lo in range(0, npart, slab):
        hi = min(npart, lo + slab)
        ctx.route_load(gen(nside, lo, hi)[0], lo, append=lo > 0)
    print(f"generated {npart} particles into HBM in {time.perf_counter() - t0:.1f} s", flush=True)
    for r in range(reps):
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.tree_build_resident(maxleaf, bdl, bdr, 0)          # from the order the previous step left, like the reference
        ctx.forces_local(theta, rcut, box, 0.5 * (bdr + bdl), bdr - bdl)
        ctx.synchronize()
        t1 = time.perf_counter()
        info, st = ctx.tree_info(), ctx.step_timings()
        nt, npairs = ctx.accumulated_counts()
        print(f"resident step {r}: {1e3 * (t1 - t0):.1f} ms | build {st['build_ms']:.1f} walk {st['walk_ms']:.1f} csr {st['csr_ms']:.1f} force {st['force_ms']:.1f} "
              f"chunks {st['chunks']} | {info['nleaf']} leaves {info['nlevel']} levels {nt} tasks {npairs} pairs "
              f"({npairs / st['force_ms'] / 1e9:.4f} Tpair/s in the force kernel, {npairs / (t1 - t0) / 1e12:.4f} Tpair/s whole step)", flush=True)
