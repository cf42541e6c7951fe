"""Multi-rank device path (p2p_b200/dist_device.py): every rank builds its tree and walks it against the trees of
all ranks on the GPU; per-rank task and pair counts must equal the oracle's restatement of the reference flow
(itself pinned against the reference run with P ranks) and the forces must agree with the fp64 oracle.
The ranks share GPU 0 and exchange through gloo here; on a multi-GPU box the same code runs over NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp
from conftest import DEMO_BOX, DEMO_MASS, DEMO_NSIDE, ROOT, THETA

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, pos, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p2p_b200
    from p2p_b200 import dist as pdist
    from p2p_b200 import dist_device
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    lp, lidx, tcenter, twidth, direct, dom = pdist.decompose(pos, DEMO_BOX, None)
    c, w = tcenter[dom], twidth[dom]
    ctx = p2p_b200.P2PContext(0)
    tm = {}
    acc, ntask, npairs = dist_device.run_device_step(ctx, lp, pos.shape[0], DEMO_BOX, 16, DEMO_NSIDE, DEMO_MASS, c - 0.5 * w,
                                                     c + 0.5 * w, int(direct[dom]), THETA, periodic=True, truncated=True, timings=tm)
    dup = ctx.csr_duplicates()
    q.put((rank, lidx, acc, ntask, npairs, dup))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_device_multirank_matches_oracle(demo_pos, world):
    import flow
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, demo_pos, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=600)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = flow.short_range_lists(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, world, True, literal_d6=False)
    acc = np.zeros((len(demo_pos), 3))
    for r in range(world):
        lidx, a, ntask, npairs, dup = got[r]
        T = ref[r]["tree"]
        want_tasks = len(ref[r]["local"][0]) + sum(len(x["tt"]) for x in ref[r]["remote"])
        want_pairs = int((T.leaf_npart[ref[r]["local"][0]].astype(np.int64) * T.leaf_npart[ref[r]["local"][1]]).sum())
        for x in ref[r]["remote"]:
            want_pairs += int((T.leaf_npart[x["tt"]].astype(np.int64) * x["image"]["npart"][x["ts"]]).sum())
        assert (ntask, npairs) == (want_tasks, want_pairs)
        assert dup == 0
        acc[lidx] = a
    want, _, _ = flow.reference_forces(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, DEMO_MASS, world, True)
    absr, _, _ = flow.reference_forces(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, DEMO_MASS, world, True, absterms=True)
    d = np.linalg.norm(acc - want, axis=1)
    na = np.linalg.norm(want, axis=1)
    e1 = (d / np.maximum(na, na.mean())).max()
    e2 = (d / np.maximum(np.linalg.norm(absr, axis=1), 1e-300)).max()
    assert e1 < 1e-5 and e2 < 1e-5, (e1, e2)


def _route_worker(rank, world, port, pos, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p2p_b200
    from p2p_b200 import dist as pdist
    from p2p_b200 import dist_device, host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    n = pos.shape[0]
    # the host routing (bit-exact with the reference's domain_decomposition) and the host tree built from it
    lp, lidx, tcenter, twidth, direct, dom = pdist.decompose(pos, DEMO_BOX, None)
    c, w = tcenter[dom], twidth[dom]
    T = host.LocalTree(lp, 16, c - 0.5 * w, c + 0.5 * w, int(direct[dom]))
    # the device routing from the same initial slab
    lo, hi = n * rank // world, n * (rank + 1) // world
    split = host.domain_setup(world, DEMO_BOX)[0]
    ctx = p2p_b200.P2PContext(0)
    acc, idx, ntask, npairs = dist_device.route_and_step(ctx, pos[lo:hi].copy(), lo, n, DEMO_BOX, 16, DEMO_NSIDE, DEMO_MASS, split, THETA)
    D = ctx.tree_download()
    same_tree = bool(np.array_equal(D["pos"], T.pos) and np.array_equal(idx, lidx[T.perm]) and
                     np.array_equal(D["leaf_ipart"], T.leaf_ipart) and np.array_equal(D["node_split"], T.node_split))
    q.put((rank, idx, acc, ntask, npairs, same_tree))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_device_routing_then_step(demo_pos, world):
    """slab -> device partition by the rank kd-tree -> exchange -> resident tree build -> lists, halo, forces: the
    routed particle ORDER must be the reference's (the tree built from it is compared bit for bit with the host
    path's, which is pinned against the reference's domain_decomposition), and the forces must match the oracle"""
    import flow
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_route_worker, args=(r, world, port, demo_pos, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=600)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acc = np.zeros((len(demo_pos), 3))
    seen = 0
    for r in range(world):
        idx, a, ntask, npairs, same_tree = got[r]
        assert same_tree
        acc[idx] = a
        seen += len(idx)
    assert seen == len(demo_pos)
    want, ntask_all, npairs_all = flow.reference_forces(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, DEMO_MASS, world, True)
    assert sum(got[r][2] for r in range(world)) == ntask_all and sum(got[r][3] for r in range(world)) == npairs_all
    d = np.linalg.norm(acc - want, axis=1)
    na = np.linalg.norm(want, axis=1)
    assert (d / np.maximum(na, na.mean())).max() < 1e-5


def _mid_worker(rank, world, port, pos, maxleaf, theta, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p2p_b200
    from p2p_b200 import dist as pdist
    from p2p_b200 import dist_device
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    lp, lidx, tcenter, twidth, direct, dom = pdist.decompose(pos, DEMO_BOX, None)
    c, w = tcenter[dom], twidth[dom]
    ctx = p2p_b200.P2PContext(0)
    acc, ntask, npairs = dist_device.run_device_step(ctx, lp, pos.shape[0], DEMO_BOX, maxleaf, DEMO_NSIDE, 1.0, c - 0.5 * w, c + 0.5 * w,
                                                     int(direct[dom]), theta, periodic=True, truncated=True, midfield=True,
                                                     literal_d6=True, p2p=False)
    q.put((rank, lidx, acc))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_device_multirank_midfield_matches_the_reference(demo_pos):
    """two ranks: multipoles all-gathered, M2L against the other rank's tree and all images; against the reference's
    own operators run as two ranks (P2P stubbed to zero, zero-shift self exchange replayed)"""
    import refrun
    if not os.path.isfile(refrun.REF_BIN):
        pytest.skip("oracle/_ref/ref_lists not built")
    world, maxleaf, theta = 2, 2, 1.0
    pos = demo_pos[::4].copy()
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mid_worker, args=(r, world, port, pos, maxleaf, theta, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((item[0], item[1:]) for item in (q.get(timeout=600) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = refrun.run(pos, DEMO_BOX, maxleaf, DEMO_NSIDE, theta, True, world)
    acc = np.zeros((len(pos), 3))
    want = np.zeros((len(pos), 3))
    for r in range(world):
        lidx, a = got[r]
        acc[lidx] = a
        want[ref[r]["part_orig_index"]] = ref[r]["acc_mid"].reshape(-1, 3)
    assert np.abs(want).max() > 0
    assert np.linalg.norm(acc - want, axis=1).max() < 1e-10 * np.linalg.norm(want, axis=1).mean()


def _resident_worker(rank, world, port, pos, vel0, q):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import p2p_b200
    from p2p_b200 import dist as pdist
    from p2p_b200 import dist_device, host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    n, maxleaf, nstep = pos.shape[0], 16, 3
    lo, hi = n * rank // world, n * (rank + 1) // world
    cell = DEMO_BOX / DEMO_NSIDE
    dkh, dd = 7.5e4 * cell, 0.4               # mean |a| of the demo is 6.7e-6: a kick adds ~0.5 cell per unit time
    split = host.domain_setup(world, DEMO_BOX)[0]
    center, width, direct = host.domain_boxes(world, DEMO_BOX, split)
    dom = host.domain_of_rank(world, rank)
    bdl, bdr = center[dom] - 0.5 * width[dom], center[dom] + 0.5 * width[dom]
    # ---- resident path: nothing returns to the host between the steps
    a = p2p_b200.P2PContext(0)
    R = dist_device.ResidentRun(a, n, DEMO_BOX, maxleaf, DEMO_NSIDE, DEMO_MASS, THETA)
    R.load(pos[lo:hi].copy(), vel0[lo:hi].copy(), lo)
    counts = [R.step(dkh, dd) for _ in range(nstep)]
    pa, va, ia = R.download()
    # ---- host mirror around the one-step device path: host routing (pinned bit-exactly to the reference's
    # domain_decomposition), exchange over gloo, forces, numpy kick / drift / wrap, arrays carried in tree order
    b = p2p_b200.P2PContext(0)
    p, v, ids = pos[lo:hi].copy(), vel0[lo:hi].copy(), np.arange(lo, hi, dtype=np.int64)
    mirror_counts = []
    for _ in range(nstep):
        lidx = np.arange(len(p), dtype=np.int64)
        send = host.domain_route(world, split, p, lidx)
        v, ids = v[lidx], ids[lidx]
        off = np.concatenate([[0], np.cumsum(send)])
        rc = pdist._a2a_counts(send.reshape(world, 1), None)[:, 0]
        p = np.concatenate(pdist._a2a_v([p[off[d]:off[d + 1]] for d in range(world)], rc, 3, np.float64, None))
        v = np.concatenate(pdist._a2a_v([v[off[d]:off[d + 1]] for d in range(world)], rc, 3, np.float64, None))
        ids = np.concatenate(pdist._a2a_v([ids[off[d]:off[d + 1]].reshape(-1, 1) for d in range(world)], rc, 1, np.int64, None))[:, 0]
        acc, nt, npr = dist_device.run_device_step(b, p, n, DEMO_BOX, maxleaf, DEMO_NSIDE, DEMO_MASS, bdl, bdr, int(direct[dom]), THETA)
        mirror_counts.append((nt, npr))
        D = b.tree_download()
        perm = D["perm"]
        p, v, ids = D["pos"], v[perm] + acc[perm] * dkh, ids[perm]
        p = p + v * dd
        for k in range(3):
            col = p[:, k]
            while (col < 0.0).any():
                col[col < 0.0] += DEMO_BOX
            while (col >= DEMO_BOX).any():
                col[col >= DEMO_BOX] -= DEMO_BOX
    moved = float(np.abs(pa - pos[ia]).max())
    q.put((rank, bool(np.array_equal(ia, ids)), bool(np.array_equal(va, v)), bool(np.array_equal(pa, p)), counts == mirror_counts, ia, moved))
    dist.barrier()
    a.close()
    b.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multirank_resident_steps_equal_host_mirror(demo_pos, world):
    """positions, velocities and ids resident across three steps on every rank, migration as device buffers: bit for bit
    the host mirror (host routing + one-step device path + numpy integrator), every particle owned exactly once"""
    rng = np.random.default_rng(6)
    vel0 = rng.normal(0.0, 0.2 * DEMO_BOX / DEMO_NSIDE, demo_pos.shape)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_resident_worker, args=(r, world, port, demo_pos, vel0, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_ids, same_vel, same_pos, same_counts, ia, moved in got:
        assert same_ids and same_vel and same_pos and same_counts, (rank, same_ids, same_vel, same_pos, same_counts)
        assert moved > 0.05 * DEMO_BOX / DEMO_NSIDE
    assert sorted(np.concatenate([g[5] for g in got]).tolist()) == list(range(len(demo_pos)))
