"""Multi-rank short-range step with the list producers ON THE DEVICES (csrc/device_tree.cuh, csrc/halo.cuh).

The reference ships 27 x P pruned halo images around a ring and re-walks each of them
(1_Indexing/src/fmm.c:1026-1145, 1_Indexing/src/remotes.c:740-809), strictly one after the other.  Here every rank

  local phase (needs no communication; compute stream)
    1. builds its kd-tree on its GPU,
    2. walks it against itself and its own periodic images, packs that list and launches the force kernel on it;
  remote phase (second, high-priority stream; hidden behind the local force kernel)
    3. copies its tree TOPOLOGY (kd cells, tight leaf bounds, sons, leaf sizes) into one block and all-gathers the blocks
       (ONE collective, no particles),
    4. walks its tree against every other rank's tree for all 27 displacements on the GPU, evaluating the sender-side
       cuts of prepare_sendtree2 on the fly (p2p_tree_walk_peers_packed) -- the task multiset is the reference's,
    5. asks each owner only for the leaves its list references (all-to-all of one byte per leaf), and
    6. receives exactly those particles (all-to-all-v of fixed-point int4; the marks, counts, offsets and the gather run
       as kernels, the host only learns the split sizes),
    7. packs the remote list into the context's second list set and runs the force kernel on it after the local one.

The local force kernel runs in its non-persistent form (p2p_set_force_blocks), so NCCL's kernels and the small kernels of
the remote phase get SM slots while it computes.  NCCL over NVLink on the GPUs; with the gloo backend (CPU test rigs,
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
    out = _all_gather_blocks(buf, P, group)
    return [out[p * mx:p * mx + int(sizes[p])] for p in range(P)]


def _all_gather_blocks(mine, P, group):
    """equal-sized 1-D blocks of every rank, concatenated in rank order (one collective)"""
    if _nccl(group):
        out = torch.empty(P * mine.numel(), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out
    parts = [torch.empty(mine.numel(), dtype=mine.dtype) for _ in range(P)]
    dist.all_gather(parts, mine.cpu(), group=group)
    return torch.cat(parts).to(mine.device)


def _all_to_all_v(send, in_split, out_split, group):
    """1-D device tensor split by in_split -> 1-D device tensor of sum(out_split) elements."""
    if _nccl(group):
        out = torch.empty(int(sum(out_split)), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, [int(x) for x in out_split], [int(x) for x in in_split], group=group)
        return out
    out = torch.empty(int(sum(out_split)), dtype=send.dtype)
    dist.all_to_all_single(out, send.cpu(), [int(x) for x in out_split], [int(x) for x in in_split], group=group)
    return out.to(send.device)


def _gather_host_ints(vals, group, dev):
    """a few host integers of every rank -> int64 array [P, len(vals)] on the host (one tiny collective)"""
    P = dist.get_world_size(group)
    mine = torch.tensor([int(v) for v in vals], dtype=torch.int64)
    if _nccl(group):
        out = torch.empty(P * len(vals), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, mine.to(dev), group=group)
        return out.cpu().numpy().reshape(P, len(vals))
    parts = [torch.empty(len(vals), dtype=torch.int64) for _ in range(P)]
    dist.all_gather(parts, mine, group=group)
    return torch.stack(parts).numpy()


class Streams:
    """The two streams of a rank's step (created once per context): compute (tree build, local walk, packing, force
    kernels) and comm (exchanges, remote walk and packing; higher priority)."""

    def __init__(self, ctx):
        dev = torch.device("cuda", ctx.device)
        self.dev = dev
        if ctx.stream_ptr:
            self.main = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
        else:
            self.main = torch.cuda.Stream(device=dev)
            ctx.set_stream(self.main.cuda_stream)
        self.comm = torch.cuda.Stream(device=dev, priority=-1)
        self.main_ptr = ctx.stream_ptr
        self.ghost_guess = 0


def _streams_of(ctx):
    s = getattr(ctx, "_dist_streams", None)
    if s is None or s.main_ptr != ctx.stream_ptr:
        s = Streams(ctx)
        ctx._dist_streams = s
    return s


def run_device_step(ctx, local_pos, npart_total, box, maxleaf, nside, mass, bdl, bdr, direct_start, theta=0.4, periodic=True,
                    truncated=True, group=None, acc_out=None, timings=None, midfield=False, literal_d6=False, p2p=True, overlap=True):
    """One rank's part of the step.  local_pos: this rank's particles (host, float64, caller's order; pinned for full
    PCIe rate); bdl/bdr/direct_start: its domain box and first split direction (host.domain_setup).
    midfield: also the multipole part (P2M/M2M on every rank, all-gather of the multipoles, M2L/L2L/L2P); literal_d6 and
    p2p=False are test knobs (replay the reference's zero-shift self exchange; skip the P2P forces).
    Returns (acc in the order of local_pos, ntask, npairs)."""
    S = _streams_of(ctx)
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, npart_total)
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    bdl, bdr = np.asarray(bdl, np.float64), np.asarray(bdr, np.float64)
    with torch.cuda.stream(S.main):
        ctx.midfield_enable(midfield, literal_d6)
        ctx.tree_build(local_pos, maxleaf, bdl, bdr, direct_start)
        ntask, npairs = lists_halo_forces(ctx, rcut, box, bdl, bdr, theta, periodic, group, timings, midfield, p2p, overlap, t0)
        acc = ctx.download_acc_original(acc_out)
    if timings is not None:
        timings["total_s"] = time.perf_counter() - t0
    return acc, ntask, npairs


def route_and_step(ctx, slab_pos, first_index, npart_total, box, maxleaf, nside, mass, split, theta=0.4, periodic=True, truncated=True,
                   group=None, timings=None, pinned_out=None, overlap=True):
    """The same step starting one stage earlier: `slab_pos` is the slab of the global particle array this rank happens to
    hold (global ids first_index ...), `split` the rank kd-tree (host.domain_setup / host.domain_relax).  The slab is
    partitioned on the device exactly as the reference's prepare_body_inOrderOf_domain does, the groups are exchanged as
    device buffers (all-to-all-v), and the tree is built from what arrived (1_Indexing/src/domains.c:163-377).
    pinned_out: optional dict reused between steps for pinned result buffers (grown on demand).
    Returns (acc in TREE order, global ids of the tree positions, ntask, npairs)."""
    P, me = dist.get_world_size(group), dist.get_rank(group)
    S = _streams_of(ctx)
    dev = S.dev
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, npart_total)
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    center, width, direct = host.domain_boxes(P, box, split)
    dom = host.domain_of_rank(P, me)
    bdl, bdr = center[dom] - 0.5 * width[dom], center[dom] + 0.5 * width[dom]
    with torch.cuda.stream(S.main):
        ctx.route_load(slab_pos, first_index)
        nloc = migrate(ctx, P, split, group, dev)
        t_route = time.perf_counter() - t0
        ctx.tree_build_resident(maxleaf, bdl, bdr, int(direct[dom]))
        ntask, npairs = lists_halo_forces(ctx, rcut, box, bdl, bdr, theta, periodic, group, timings, False, True, overlap, t0)
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


def migrate(ctx, P, split, group, dev):
    """The particles resident on this rank's GPU (p2p_route_load, or the previous step's after the drift) move to the ranks
    that own them under the rank kd-tree `split`: device partition with the reference's order, all-to-all-v of x, y, z and
    the global id as device buffers, import in source-rank order (1_Indexing/src/domains.c:163-377).  Returns the new count."""
    n = ctx.resident_count()
    send = ctx.route_partition(P, split)
    if P == 1:
        return n
    bufs = [torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(max(n, 1), dtype=torch.int32, device=dev)]
    ctx.route_export(*[b.data_ptr() for b in bufs])
    recv = _gather_host_ints(send, group, dev)[:, dist.get_rank(group)]
    got = [_all_to_all_v(b[:n], [int(v) for v in send], [int(v) for v in recv], group) for b in bufs]
    nloc = int(recv.sum())
    ctx.route_import(*[g.data_ptr() for g in got], nloc)
    return nloc


def lists_halo_forces(ctx, rcut, box, bdl, bdr, theta, periodic, group, timings=None, midfield=False, p2p=True, overlap=True, t0=None):
    """after the tree build: local phase (own tree and images) and remote phase (topology all-gather, walk against every
    other rank's tree, halo fetch), packing and forces.  Returns (ntask, npairs) of this rank."""
    P, me = dist.get_world_size(group), dist.get_rank(group)
    S = _streams_of(ctx)
    A, B, dev = S.main, (S.comm if overlap else S.main), S.dev
    t0 = time.perf_counter() if t0 is None else t0
    period = box if periodic else 0.0
    tc, tw = 0.5 * (bdr + bdl), bdr - bdl
    info = ctx.tree_info()
    nl, nn = info["nleaf"], info["nnode"]
    ctx.set_rank(me, P)
    t1 = time.perf_counter()
    ev = {k: torch.cuda.Event(enable_timing=True) for k in ("c0", "c1", "c2", "c3", "w0", "w1", "x0", "x1")}
    if P > 1:
        sizes = _gather_host_ints([nl, nn], group, dev)
        nls, nns = [int(v) for v in sizes[:, 0]], [int(v) for v in sizes[:, 1]]
        nlmax, nnmax = max(nls), max(nns)
        G = sum(nls) - nl
        ctx.reserve_ghosts(G, max(S.ghost_guess, ctx.npart // 8))
    built = torch.cuda.Event()
    built.record(A)
    # ---- local phase: my tree against itself and its own periodic images (walk, packing, force kernel; in target chunks
    # when the list would not fit)
    nvtx = torch.cuda.nvtx
    nvtx.range_push("p2p/local phase: walk + pack + forces")
    ctx.set_force_blocks(4 if (P > 1 and overlap) else 0)        # warps retire after ~0.3 ms
    ctx.forces_local(theta, rcut, period, tc, tw, p2p)
    ctx.set_force_blocks(0)
    nvtx.range_pop()
    t2 = time.perf_counter()
    local_done = torch.cuda.Event()
    local_done.record(A)
    nbody = 0
    topo_bytes = 0
    remote_walk_ms = 0.0
    if P > 1:
        # ---- remote phase on the comm stream, while the local force kernel runs
        B.wait_event(built)
        with torch.cuda.stream(B):
            ctx.set_stream(B.cuda_stream)
            ctx.swap_lists()
            ctx.clear_tasks()
            ev["c0"].record(B)
            nvtx.range_push("p2p/remote phase: topology all-gather")
            stride = ctx.topology_stride(nlmax, nnmax)
            mine = torch.empty(stride, dtype=torch.uint8, device=dev)
            ctx.tree_export_packed(mine.data_ptr(), nlmax, nnmax)
            topo = _all_gather_blocks(mine, P, group)
            topo_bytes = int(topo.numel())
            if midfield:
                M_mine = torch.zeros((nlmax + nnmax) * 20, dtype=torch.float64, device=dev)
                ctx.midfield_multipoles(M_mine.data_ptr())
                M_all = _all_gather_blocks(M_mine, P, group)
            ev["c1"].record(B)
            nvtx.range_pop()
            nvtx.range_push("p2p/remote phase: walk against the peers' trees")
            ev["w0"].record(B)
            ctx.tree_walk_peers_packed(theta, rcut, period, tc, tw, me, nls, nns, topo.data_ptr(), nlmax, nnmax, 0)
            remote_walk_ms = ctx.tree_info()["ms_walk"]
            ev["w1"].record(B)
            nvtx.range_pop()
            nvtx.range_push("p2p/remote phase: halo plan + exchange")
            ev["c2"].record(B)
            marks = torch.empty(max(G, 1), dtype=torch.uint8, device=dev)
            need = ctx.halo_plan_need(topo.data_ptr(), me, nls, nlmax, nnmax, marks.data_ptr())
            asked = _all_to_all_v(marks[:G], [nls[p] if p != me else 0 for p in range(P)], [nl if q != me else 0 for q in range(P)], group)
            give = ctx.halo_plan_give(asked.data_ptr(), P - 1)
            others = [p for p in range(P) if p != me]
            n_give = int(give.sum())
            send = torch.empty(max(n_give, 1) * 4, dtype=torch.int32, device=dev)
            ctx.halo_gather(asked.data_ptr(), P - 1, send.data_ptr())
            in_split, out_split = [0] * P, [0] * P
            for i, p in enumerate(others):
                in_split[p], out_split[p] = 4 * int(give[i]), 4 * int(need[p])
            ghosts = _all_to_all_v(send[:4 * n_give], in_split, out_split, group)
            nbody = int(need.sum())
            ctx.halo_set_particles(ghosts.data_ptr() if nbody else None, nbody)
            S.ghost_guess = max(S.ghost_guess, int(1.25 * nbody))
            ev["c3"].record(B)
            nvtx.range_pop()
            nvtx.range_push("p2p/remote phase: pack + forces")
            ctx.build_csr()
            B.wait_event(local_done)        # both force kernels accumulate into the same accelerations
            if p2p:
                ctx.compute()
            if midfield:
                ctx.midfield_compute_peers_packed(P, topo.data_ptr(), nlmax, nnmax, M_all.data_ptr())
            remote_done = torch.cuda.Event()
            remote_done.record(B)
            nvtx.range_pop()
        A.wait_event(remote_done)
        ctx.set_stream(A.cuda_stream)
        nt_r = ctx.counts()[0]
        ms_force_r, ms_csr_r = ctx.last_timings()
        ctx.swap_lists()
    else:
        nt_r = 0
        ms_force_r = ms_csr_r = 0.0
        if midfield:
            ctx.midfield_compute()
    ntask, npairs = ctx.accumulated_counts() if p2p else (ctx.counts()[0] + nt_r, 0)
    t5 = time.perf_counter()
    if timings is not None:
        st = ctx.step_timings()
        comm_ms = (ev["c0"].elapsed_time(ev["c1"]) + ev["c2"].elapsed_time(ev["c3"])) if P > 1 else 0.0
        timings.update(build_s=t1 - t0, local_phase_s=t2 - t1, remote_phase_s=t5 - t2, build_ms=info["ms_build"], walk_ms=st["walk_ms"] + remote_walk_ms,
                       csr_ms=st["csr_ms"] + ms_csr_r, force_ms=st["force_ms"] + ms_force_r, force_local_ms=st["force_ms"], force_remote_ms=ms_force_r,
                       comm_ms=comm_ms, remote_walk_ms=remote_walk_ms, chunks=st["chunks"],
                       ghost_particles=nbody, remote_tasks=nt_r, topology_bytes=topo_bytes)
    return ntask, npairs


class ResidentRun:
    """Multi-rank device-resident stepping (SURVEY 8f N4 + row a12): positions, velocities and ids stay in HBM on every
    rank across steps.  Each step: the particles migrate to the ranks that own them after the last drift (device partition
    by the rank kd-tree, all-to-all-v of the seven state arrays as device buffers -- domain_decomposition,
    1_Indexing/src/domains.c:298-377), the tree is built from the arrived order, forces as in lists_halo_forces, then kick
    and drift with the reference's arithmetic (1_Indexing/src/photoNs.c:161-208).  Nothing crosses PCIe per step."""

    def __init__(self, ctx, npart_total, box, maxleaf, nside, mass, theta=0.4, truncated=True, midfield=False, group=None):
        self.ctx, self.group = ctx, group
        self.P, self.me = dist.get_world_size(group), dist.get_rank(group)
        self.box, self.maxleaf, self.theta, self.midfield = box, maxleaf, theta, midfield
        self.rs, self.rcut, self.eps = host.derived_params(box, nside, npart_total)
        ctx.set_physics(mass, self.eps, self.rs if truncated else 0.0)
        ctx.set_box([0.0, 0.0, 0.0], box)
        self.split = host.domain_setup(self.P, box)[0]
        self.S = _streams_of(ctx)

    def load(self, slab_pos, slab_vel, first_index):
        """the slab of the global arrays this rank starts from (global ids first_index ...)"""
        with torch.cuda.stream(self.S.main):
            self.ctx.resident_load(slab_pos, slab_vel, first_index)

    def _migrate(self):
        ctx, P, dev = self.ctx, self.P, self.S.dev
        n = ctx.resident_count()
        send = ctx.resident_partition(P, self.split)
        if P == 1:
            return n
        bufs = [torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(6)]
        ids = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ctx.resident_export([b.data_ptr() for b in bufs], ids.data_ptr())
        recv = _gather_host_ints(send, self.group, dev)[:, self.me]
        s, r = [int(v) for v in send], [int(v) for v in recv]
        got = [_all_to_all_v(b[:n], s, r, self.group) for b in bufs]
        gid = _all_to_all_v(ids[:n], s, r, self.group)
        nloc = int(recv.sum())
        ctx.resident_import([g.data_ptr() for g in got], gid.data_ptr(), nloc)
        return nloc

    def step(self, dkh, dd, timings=None, overlap=True):
        """migrate -> build -> forces -> kick (vel += acc dkh) -> drift (pos += vel dd, wrapped).  Returns (ntask, npairs)."""
        ctx = self.ctx
        with torch.cuda.stream(self.S.main):
            self._migrate()
            center, width, direct = host.domain_boxes(self.P, self.box, self.split)
            dom = host.domain_of_rank(self.P, self.me)
            bdl, bdr = center[dom] - 0.5 * width[dom], center[dom] + 0.5 * width[dom]
            ctx.midfield_enable(self.midfield, False)
            ctx.resident_build(self.maxleaf, bdl, bdr, int(direct[dom]))
            out = lists_halo_forces(ctx, self.rcut, self.box, bdl, bdr, self.theta, True, self.group, timings, self.midfield, True, overlap)
            ctx.resident_kick(dkh)
            ctx.resident_drift(dd, self.box)
        return out

    def relax(self, work):
        """the reference's work-weighted split relaxation (1_Indexing/src/photoNs.c:295-306); takes effect at the next migration"""
        w_all = _gather_host_ints([int(work)], self.group, self.S.dev)[:, 0].astype(np.float64)
        self.split = host.domain_relax(self.P, self.box, self.split, w_all)

    def download(self):
        """(positions, velocities, ids) of this rank's particles, in the resident order"""
        with torch.cuda.stream(self.S.main):
            return self.ctx.resident_download()
