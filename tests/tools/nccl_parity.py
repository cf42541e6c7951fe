"""Force and list parity of the multi-rank device path over NCCL (one rank per GPU; run under torchrun):
  torchrun --nproc-per-node N tests/tools/nccl_parity.py [out.json]
Every rank starts from its slab of the demo IC, the slabs are routed on the devices (all-to-all-v of device buffers), trees
are built, topology all-gathered, halos fetched leaf-granular -- all through the NCCL branches of p2p_b200/dist_device.py
that the gloo test rigs (several ranks sharing one GPU) cannot execute.  Rank 0 compares per-rank task / pair counts and
the accelerations with the fp64 oracle's restatement of the reference flow at the same number of ranks, and checks that the
step with the halo overlap gives bit-identical accelerations to the step without."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")]
import p2p_b200  # noqa: E402
from p2p_b200 import dist_device, host  # noqa: E402

BOX, NSIDE, THETA, MASS, MAXLEAF = 100000.0, 32, 0.4, 211.75382579190332, 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_lcdm_pos_f32.npy")).astype(np.float64)
n = len(pos)
lo, hi = n * rank // world, n * (rank + 1) // world
split = host.domain_setup(world, BOX)[0]
ctx = p2p_b200.P2PContext(local)
res = {}
for overlap in (True, False):
    acc, idx, ntask, npairs = dist_device.route_and_step(ctx, pos[lo:hi].copy(), lo, n, BOX, MAXLEAF, NSIDE, MASS, split, THETA, overlap=overlap)
    res[overlap] = (acc.copy(), idx.copy(), ntask, npairs)
same = bool(np.array_equal(res[True][0], res[False][0]) and np.array_equal(res[True][1], res[False][1]))
acc, idx, ntask, npairs = res[True]
out = [None] * world
dist.all_gather_object(out, (rank, idx, acc, ntask, npairs, same))
verdict = {"world": world, "backend": "nccl"}
if rank == 0:
    import flow
    ref = flow.short_range_lists(pos, BOX, MAXLEAF, NSIDE, THETA, world, True, literal_d6=False)
    got = np.zeros((n, 3))
    counts_ok, overlap_ok = True, True
    for r, idx_r, a, nt, npr, sm in out:
        T = ref[r]["tree"]
        want_tasks = len(ref[r]["local"][0]) + sum(len(x["tt"]) for x in ref[r]["remote"])
        want_pairs = int((T.leaf_npart[ref[r]["local"][0]].astype(np.int64) * T.leaf_npart[ref[r]["local"][1]]).sum())
        for x in ref[r]["remote"]:
            want_pairs += int((T.leaf_npart[x["tt"]].astype(np.int64) * x["image"]["npart"][x["ts"]]).sum())
        counts_ok &= (nt, npr) == (want_tasks, want_pairs)
        overlap_ok &= sm
        got[idx_r] = a
    want, _, _ = flow.reference_forces(pos, BOX, MAXLEAF, NSIDE, THETA, MASS, world, True)
    absr, _, _ = flow.reference_forces(pos, BOX, MAXLEAF, NSIDE, THETA, MASS, world, True, absterms=True)
    d = np.linalg.norm(got - want, axis=1)
    na = np.linalg.norm(want, axis=1)
    verdict.update(counts_equal_oracle=bool(counts_ok), overlap_bit_identical=bool(overlap_ok),
                   e1=float((d / np.maximum(na, na.mean())).max()), e2=float((d / np.maximum(np.linalg.norm(absr, axis=1), 1e-300)).max()),
                   tasks=int(sum(o[3] for o in out)), pairs=int(sum(o[4] for o in out)))
    verdict["ok"] = bool(counts_ok and overlap_ok and verdict["e1"] < 1e-5 and verdict["e2"] < 1e-5)
    print(json.dumps(verdict), flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(verdict, f)
dist.barrier()
ctx.close()
dist.destroy_process_group()
