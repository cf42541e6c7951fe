"""Multi-GPU device-resident short-range step (p2p_b200/dist_device.py) under torchrun, without the host-list legs of
bench.py: generate, route, then time the step (max over ranks, wall clock between barriers).
usage: torchrun --nproc-per-node N tools/dist_device_step.py [nside] [maxleaf] [reps] [--clustered] [--relax=K]
--route: time the step from the slab of the global array each rank holds (device routing + exchange + resident build)
instead of from already routed particles.
--relax=K: K further steps, each after the reference's work-weighted split relaxation (p2p_domain_relax, fed with the
per-rank task counts as 1_Indexing/src/photoNs.c:295-306 does) and a re-routing of the particles."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200  # noqa: E402
from p2p_b200 import dist as pdist, dist_device, host, synth  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nside = int(args[0]) if len(args) > 0 else 256
maxleaf = int(args[1]) if len(args) > 1 else 32
reps = int(args[2]) if len(args) > 2 else 4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, box = (synth.clustered(nside) if "--clustered" in sys.argv else synth.zeldovich_like(nside))
npart_total = pos.shape[0]
relax = max([int(a.split("=")[1]) for a in sys.argv if a.startswith("--relax=")] + [0])
split = host.domain_setup(world, box)[0]
ctx = p2p_b200.P2PContext(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
res = []
history = []
for r in range(reps + relax):
    if r == 0 or r >= reps:
        if r >= reps:                            # feedback: splits move towards equal work
            w_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(w_all, torch.tensor([float(ntask)], dtype=torch.float64, device="cuda"))
            split = host.domain_relax(world, box, split, np.array([float(x.item()) for x in w_all]))
        lp, lidx, tcenter, twidth, direct, dom = pdist.decompose(pos, box, None, split)
        c, w = tcenter[dom], twidth[dom]
        lp_t = torch.from_numpy(np.ascontiguousarray(lp)).pin_memory()
        acc_t = torch.empty((lp.shape[0], 3), dtype=torch.float64).pin_memory()
    tm = {}
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, ntask, npairs = dist_device.run_device_step(ctx, lp_t.numpy(), npart_total, box, maxleaf, nside, 1.0, c - 0.5 * w, c + 0.5 * w,
                                                   int(direct[dom]), 0.4, periodic=True, truncated=True, acc_out=acc_t.numpy(), timings=tm)
    dist.barrier(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tot = torch.tensor([ntask, npairs, lp.shape[0]], dtype=torch.float64, device="cuda")
    mx = tot.clone()
    dist.all_reduce(tot); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    res.append((dt, tm, tot.tolist(), mx.tolist()))
    history.append({"step_s": dt, "imbalance_pairs": 1.0 - tot.tolist()[1] / (world * mx.tolist()[1]),
                    "imbalance_tasks": 1.0 - tot.tolist()[0] / (world * mx.tolist()[0])})
routed = None
if "--route" in sys.argv:
    n = pos.shape[0]
    lo, hi = n * rank // world, n * (rank + 1) // world
    slab = torch.from_numpy(np.ascontiguousarray(pos[lo:hi])).pin_memory()
    pinned = {}
    for r in range(3):
        tm2 = {}
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, nt2, np2 = dist_device.route_and_step(ctx, slab.numpy(), lo, npart_total, box, maxleaf, nside, 1.0, split, 0.4, timings=tm2, pinned_out=pinned)
        dist.barrier(); torch.cuda.synchronize()
        routed = (time.perf_counter() - t0, tm2)
if rank == 0:
    dt, tm, tot, mx = res[-1]
    print(json.dumps({"n_gpus": world, "nside": nside, "maxleaf": maxleaf, "step_s": dt, "all_reps_s": [r[0] for r in res], "pairs": int(tot[1]),
                      "tasks": int(tot[0]), "pair_per_s_whole_step": tot[1] / dt, "imbalance": 1.0 - tot[1] / (world * mx[1]),
                      "rank0_breakdown": tm,
                      "relaxation_history": history[reps - 1:] if relax else None,
                      "routed_step_s": routed[0] if routed else None, "routed_rank0_breakdown": routed[1] if routed else None}), flush=True)
dist.barrier()
dist.destroy_process_group()
