"""Multi-rank short-range step with the list producers ON THE DEVICES (csrc/device_tree.cuh).

The reference ships 27 x P pruned halo images around a ring and re-walks each of them
(1_Indexing/src/fmm.c:1026-1145, 1_Indexing/src/remotes.c:740-809).  Here every rank
  1. builds its kd-tree on its GPU (p2p_tree_build),
  2. all-gathers the tree TOPOLOGY of all ranks (kd cells + sons + leaf sizes: a few MB, no particles),
  3. walks its tree against every rank's tree for all 27 displacements on the GPU, evaluating the sender-side
     cuts of prepare_sendtree2 on the fly (p2p_tree_walk_peers) -- the task multiset is the reference's,
  4. asks each owner only for the leaves its list references (all-to-all of one byte per leaf), and
  5. receives exactly those particles (all-to-all-v of fixed-point int4), already in the global periodic frame,
then packs the list and runs the force kernel.  NCCL over NVLink on the GPUs; with the gloo backend (CPU test rigs,
several ranks sharing one GPU) the same exchanges are staged through host memory."""
import time

import numpy as np
import torch
import torch.distributed as dist

from . import host


def _nccl(group):
    return dist.get_backend(group) == "nccl"


def _all_gather_v(t, sizes, group):
    """t: 1-D device tensor of this rank (length sizes[me]) -> list of 1-D device tensors, one per rank."""
    P, mx = len(sizes), int(max(sizes))
    buf = torch.zeros(mx, dtype=t.dtype, device=t.device)
    buf[:t.numel()] = t
    if _nccl(group):
        out = torch.empty(P * mx, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        parts = [torch.empty(mx, dtype=t.dtype) for _ in range(P)]
        dist.all_gather(parts, buf.cpu(), group=group)
        out = torch.cat(parts).to(t.device)
    return [out[p * mx:p * mx + int(sizes[p])] for p in range(P)]


def _all_to_all_v(send, in_split, out_split, group):
    """1-D device tensor split by in_split -> 1-D device tensor of sum(out_split) elements."""
    if _nccl(group):
        out = torch.empty(int(sum(out_split)), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, [int(x) for x in out_split], [int(x) for x in in_split], group=group)
        return out
    out = torch.empty(int(sum(out_split)), dtype=send.dtype)
    dist.all_to_all_single(out, send.cpu(), [int(x) for x in out_split], [int(x) for x in in_split], group=group)
    return out.to(send.device)


def _stream_of(ctx, dev):
    if not ctx.stream_ptr:                      # torch ops, collectives and the library's kernels share one stream
        ctx._torch_stream = torch.cuda.Stream(device=dev)
        ctx.set_stream(ctx._torch_stream.cuda_stream)
    return torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)


def run_device_step(ctx, local_pos, npart_total, box, maxleaf, nside, mass, bdl, bdr, direct_start, theta=0.4, periodic=True,
                    truncated=True, group=None, acc_out=None, timings=None, midfield=False, literal_d6=False, p2p=True):
    """One rank's part of the step.  local_pos: this rank's particles (host, float64, caller's order; pinned for full
    PCIe rate); bdl/bdr/direct_start: its domain box and first split direction (host.domain_setup).
    midfield: also the multipole part (P2M/M2M on every rank, all-gather of the multipoles, M2L/L2L/L2P); literal_d6 and
    p2p=False are test knobs (replay the reference's zero-shift self exchange; skip the P2P forces).
    Returns (acc in the order of local_pos, ntask, npairs)."""
    dev = torch.device("cuda", ctx.device)
    stream = _stream_of(ctx, dev)
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, npart_total)
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    bdl, bdr = np.asarray(bdl, np.float64), np.asarray(bdr, np.float64)
    with torch.cuda.stream(stream):
        ctx.midfield_enable(midfield, literal_d6)
        ctx.tree_build(local_pos, maxleaf, bdl, bdr, direct_start)
        ntask, npairs = _lists_halo_forces(ctx, dev, t0, rcut, box, bdl, bdr, theta, periodic, group, timings, midfield, p2p)
        acc = ctx.download_acc_original(acc_out)
    if timings is not None:
        timings["total_s"] = time.perf_counter() - t0
    return acc, ntask, npairs


def route_and_step(ctx, slab_pos, first_index, npart_total, box, maxleaf, nside, mass, split, theta=0.4, periodic=True, truncated=True,
                   group=None, timings=None, pinned_out=None):
    """The same step starting one stage earlier: `slab_pos` is the slab of the global particle array this rank happens to
    hold (global ids first_index ...), `split` the rank kd-tree (host.domain_setup / host.domain_relax).  The slab is
    partitioned on the device exactly as the reference's prepare_body_inOrderOf_domain does, the groups are exchanged as
    device buffers (all-to-all-v), and the tree is built from what arrived (1_Indexing/src/domains.c:163-377).
    pinned_out: optional dict reused between steps for pinned result buffers (grown on demand).
    Returns (acc in TREE order, global ids of the tree positions, ntask, npairs)."""
    P, me = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    stream = _stream_of(ctx, dev)
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, npart_total)
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    center, width, direct = host.domain_boxes(P, box, split)
    dom = host.domain_of_rank(P, me)
    bdl, bdr = center[dom] - 0.5 * width[dom], center[dom] + 0.5 * width[dom]
    with torch.cuda.stream(stream):
        n = slab_pos.shape[0]
        ctx.route_load(slab_pos, first_index)
        send = ctx.route_partition(P, split)
        bufs = [torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(max(n, 1), dtype=torch.int32, device=dev)]
        ctx.route_export(*[b.data_ptr() for b in bufs])
        s_t = torch.from_numpy(send.astype(np.int64))
        r_t = torch.empty(P, dtype=torch.int64)
        if _nccl(group):
            rd = torch.empty(P, dtype=torch.int64, device=dev)
            dist.all_to_all_single(rd, s_t.to(dev), group=group)
            r_t = rd.cpu()
        else:
            dist.all_to_all_single(r_t, s_t, group=group)
        recv = [int(v) for v in r_t]
        got = [_all_to_all_v(b[:n], [int(v) for v in send], recv, group) for b in bufs]
        nloc = int(sum(recv))
        ctx.route_import(*[g.data_ptr() for g in got], nloc)
        t_route = time.perf_counter() - t0
        ctx.tree_build_resident(maxleaf, bdl, bdr, int(direct[dom]))
        ntask, npairs = _lists_halo_forces(ctx, dev, t0, rcut, box, bdl, bdr, theta, periodic, group, timings)
        if pinned_out is not None:
            if pinned_out.get("n", 0) < nloc:
                pinned_out["acc"] = torch.empty((int(nloc * 1.1) + 16, 3), dtype=torch.float64).pin_memory()
                pinned_out["idx"] = torch.empty(int(nloc * 1.1) + 16, dtype=torch.int64).pin_memory()
                pinned_out["n"] = int(nloc * 1.1) + 16
            acc = ctx.download_acc(pinned_out["acc"].numpy()[:nloc])
            idx = ctx.download_index(pinned_out["idx"].numpy()[:nloc])
        else:
            acc = ctx.download_acc()
            idx = ctx.download_index()
    if timings is not None:
        timings["route_s"] = t_route
        timings["total_s"] = time.perf_counter() - t0
    return acc, idx, ntask, npairs


def _lists_halo_forces(ctx, dev, t0, rcut, box, bdl, bdr, theta, periodic, group, timings, midfield=False, p2p=True):
    """after the tree build: topology all-gather, walk against every rank's tree, halo fetch, packing, forces"""
    P, me = dist.get_world_size(group), dist.get_rank(group)
    if True:
        info = ctx.tree_info()
        nl, nn = info["nleaf"], info["nnode"]
        t1 = time.perf_counter()
        # ---- topology of every rank
        mine = torch.tensor([nl, nn, np.float64(info["max_leaf_width"]).view(np.int64)], dtype=torch.int64)
        sizes = [torch.empty(3, dtype=torch.int64) for _ in range(P)]
        if _nccl(group):
            g = [torch.empty(3, dtype=torch.int64, device=dev) for _ in range(P)]
            dist.all_gather(g, mine.to(dev), group=group)
            sizes = [x.cpu() for x in g]
        else:
            dist.all_gather(sizes, mine, group=group)
        nls = [int(s[0]) for s in sizes]
        nns = [int(s[1]) for s in sizes]
        box_t = torch.empty((nl + nn) * 6, dtype=torch.float64, device=dev)
        son_t = torch.empty(nn * 2, dtype=torch.int32, device=dev)
        leaf_t = torch.empty(nl * 2, dtype=torch.int32, device=dev)
        tb_t = torch.empty(nl * 6, dtype=torch.float64, device=dev)
        ctx.tree_export(box_t.data_ptr(), son_t.data_ptr(), leaf_t.data_ptr(), tb_t.data_ptr())
        boxes = _all_gather_v(box_t, [(a + b) * 6 for a, b in zip(nls, nns)], group)
        sons = _all_gather_v(son_t, [2 * b for b in nns], group)
        leaves = _all_gather_v(leaf_t, [2 * a for a in nls], group)
        box_all, son_all = torch.cat(boxes), torch.cat(sons)
        tb_all = torch.cat(_all_gather_v(tb_t, [6 * a for a in nls], group))
        if midfield:
            M_t = torch.empty((nl + nn) * 20, dtype=torch.float64, device=dev)
            ctx.midfield_multipoles(M_t.data_ptr())
            M_all = torch.cat(_all_gather_v(M_t, [(a + b) * 20 for a, b in zip(nls, nns)], group))
        t2 = time.perf_counter()
        # ---- lists: my tree against every rank's tree, all displacements
        ctx.clear_tasks()
        ctx.tree_walk_peers(theta, rcut, box if periodic else 0.0, 0.5 * (bdr + bdl), bdr - bdl, me, nls, nns, box_all.data_ptr(),
                            son_all.data_ptr(), tb_all.data_ptr())
        t3 = time.perf_counter()
        # ---- which remote leaves do I need; which of mine do the others need
        others = [p for p in range(P) if p != me]
        G = sum(nls[p] for p in others)
        marks = torch.zeros(max(G, 1), dtype=torch.uint8, device=dev)
        ctx.ghost_marks(marks.data_ptr())
        marks = marks[:G]
        cnt_others = torch.cat([leaves[p].view(-1, 2)[:, 1] for p in others]).to(torch.int64) if others else torch.zeros(0, dtype=torch.int64, device=dev)
        need = marks.to(torch.int64) * cnt_others                           # particles wanted per remote leaf
        asked = _all_to_all_v(marks, [nls[p] if p != me else 0 for p in range(P)], [nl if q != me else 0 for q in range(P)], group)
        cnt_me = leaves[me].view(-1, 2)[:, 1].to(torch.int64)
        give = asked.view(len(others), nl).to(torch.int64) * cnt_me[None, :] if others else torch.zeros((0, nl), dtype=torch.int64, device=dev)
        flat = give.reshape(-1)
        off = torch.cumsum(flat, 0) - flat                                  # requester-major offsets into the send buffer
        bounds = np.concatenate([[0], np.cumsum([nls[p] for p in others])]).astype(np.int64)
        totals = torch.cat([give.sum(1), torch.stack([need[bounds[i]:bounds[i + 1]].sum() for i in range(len(others))])
                            if others else torch.zeros(0, dtype=torch.int64, device=dev)]).tolist()
        n_give, n_need = totals[:len(others)], totals[len(others):]
        send = torch.empty(max(int(sum(n_give)), 1) * 4, dtype=torch.int32, device=dev)
        askedv = asked.view(len(others), nl) if others else asked
        for qi in range(len(others)):
            ctx.gather_leaves(askedv[qi].data_ptr(), off[qi * nl:(qi + 1) * nl].data_ptr(), send.data_ptr())
        in_split, out_split = [0] * P, [0] * P
        for i, p in enumerate(others):
            in_split[p], out_split[p] = 4 * int(n_give[i]), 4 * int(n_need[i])
        ghosts = _all_to_all_v(send[:sum(in_split)], in_split, out_split, group)
        start = (torch.cumsum(need, 0) - need).to(torch.int32)
        count = need.to(torch.int32)
        nbody = int(sum(n_need))
        ctx.set_ghosts_device(ghosts.data_ptr() if nbody else None, nbody, start.data_ptr() if G else None,
                              count.data_ptr() if G else None, G)
        t4 = time.perf_counter()
        # ---- pack, forces
        ctx.build_csr()
        if p2p:
            ctx.compute()
        ntask, npairs = ctx.counts()
        if midfield:
            ctx.midfield_compute_peers(nls, nns, box_all.data_ptr(), M_all.data_ptr())
    t5 = time.perf_counter()
    if timings is not None:
        ms_force, ms_csr = ctx.last_timings()
        timings.update(build_s=t1 - t0, topology_s=t2 - t1, walk_s=t3 - t2, halo_s=t4 - t3, force_s=t5 - t4,
                       build_ms=info["ms_build"], walk_ms=ctx.tree_info()["ms_walk"], csr_ms=ms_csr, force_ms=ms_force,
                       ghost_particles=nbody, ghost_leaves_referenced=int((count > 0).sum().item()) if G else 0,
                       topology_bytes=int(box_all.numel() * 8 + son_all.numel() * 4))
    return ntask, npairs
