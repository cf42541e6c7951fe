"""world_size-2 (and 4) gloo test of the multi-rank host flow: domain routing, tree build, grouped
halo exchange and remote walks through torch.distributed must reproduce the oracle's restatement of
the reference flow task for task (which itself is pinned against the reference run with P ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp
from conftest import DEMO_BOX, DEMO_NSIDE, ROOT, THETA


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, pos, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from p2p_b200 import dist as pdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = pdist.build_lists(pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, nthreads=2, literal_d6=True)
    q.put((rank, L.tree.pos, L.orig_index, L.tree.leaf_npart, L.tt, L.ts, L.remote_calls, L.gtt, L.ghost_count[L.gts],
           L.ghost_pos[L.ghost_start[L.gts[:200]]] if len(L.gts) else np.zeros((0, 3))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_flow_matches_oracle(demo_pos, world):
    import flow
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, demo_pos, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=300)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = flow.short_range_lists(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, world, True, literal_d6=True)
    for r in range(world):
        tpos, oidx, leaf_npart, tt, ts, calls, gtt, gcount, gsample = got[r]
        T = ref[r]["tree"]
        assert np.array_equal(tpos, T.pos) and np.array_equal(oidx, ref[r]["orig_index"])
        assert np.array_equal(leaf_npart, T.leaf_npart[:T.nleaf])
        assert np.array_equal(tt, ref[r]["local"][0]) and np.array_equal(ts, ref[r]["local"][1])
        assert [c[2] for c in calls] == [len(x["tt"]) for x in ref[r]["remote"]]
        assert [c[1] for c in calls] == [x["src_rank"] for x in ref[r]["remote"]]
        rtt = np.concatenate([x["tt"] for x in ref[r]["remote"]])
        rcnt = np.concatenate([x["image"]["npart"][x["ts"]] for x in ref[r]["remote"]])
        assert np.array_equal(gtt, rtt) and np.array_equal(gcount, rcnt)
        first = np.concatenate([x["image"]["body"][x["image"]["son"][x["ts"], 0]] for x in ref[r]["remote"] if len(x["tt"])])[:200]
        assert np.array_equal(gsample, first)


def _xchg_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
    import torch
    import torch.distributed as dist
    from p2p_b200 import dist_device
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # variable-length all-gather: rank r contributes r + 2 values
    mine = torch.arange(rank + 2, dtype=torch.float64) + 100.0 * rank
    parts = dist_device._all_gather_v(mine, [r + 2 for r in range(world)], None)
    ok = all(torch.equal(parts[r], torch.arange(r + 2, dtype=torch.float64) + 100.0 * r) for r in range(world))
    # all-to-all-v: rank r sends (r + p + 1) copies of the value 10 r + p to rank p
    send = torch.cat([torch.full((rank + p + 1,), 10 * rank + p, dtype=torch.int32) for p in range(world)])
    got = dist_device._all_to_all_v(send, [rank + p + 1 for p in range(world)], [q_ + rank + 1 for q_ in range(world)], None)
    want = torch.cat([torch.full((q_ + rank + 1,), 10 * q_ + rank, dtype=torch.int32) for q_ in range(world)])
    ok = ok and torch.equal(got, want)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_exchange_helpers_of_the_device_path():
    """the collectives plumbing of p2p_b200/dist_device.py (all-gather-v, all-to-all-v), world size 2 and 3, CPU tensors"""
    for world in (2, 3):
        port = _free_port()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=300) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(res[r] for r in range(world))
