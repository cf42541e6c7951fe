"""The short-range P2P step on one rank: tree -> lists (+ periodic-image / halo ghosts) -> device.

Mirrors the reference's fmm_prepare / fmm_task / fmm_ext sequence for the P2P part
(1_Indexing/src/photoNs.c:97-123, 1_Indexing/src/fmm.c:947-1024,1026-1145) with the intended
semantics: every local task once (the zero-shift self exchange of SURVEY defect D6 is not
repeated), sources of the 26 periodic images and of other domains as displaced ghost leaves."""
import time

import numpy as np

from . import host
from .binding import P2PContext

SHIFTS = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) != (0, 0, 0)]
"""The 26 periodic displacements in the order 1_Indexing/src/fmm.c:1084-1106 issues them."""


def check_wrap_condition(tree, box, rcut, what="list"):
    """The device stores fixed-point coordinates whose differences wrap to the minimal image
    (csrc/p2p_kernel.cuh).  That equals the displaced-copy semantics of the reference
    (1_Indexing/src/remotes.c:360-366) iff every separation a listed leaf pair implies stays below
    box/2 per axis.  A listed pair has a box gap below r_cut, so r_cut + 2 * (largest leaf width) is a
    bound on any such separation."""
    wmax = float(tree.leaf_width.max()) if tree.nleaf else 0.0
    if rcut + 2.0 * wmax >= 0.5 * box:
        raise ValueError(f"{what}: r_cut ({rcut:g}) + 2 x largest leaf width ({wmax:g}) reaches box/2 ({0.5 * box:g}); "
                         "the periodic box is too small for minimal-image coordinates (needs roughly NSIDE >= 32)")


class HostLists:
    """Everything the device needs for one rank, produced on the host."""

    def __init__(self):
        self.tree = None
        self.tt = self.ts = None                 # local tasks
        self.ghost_pos = np.zeros((0, 3))        # displaced ghost bodies
        self.ghost_start = np.zeros(0, np.int32)
        self.ghost_count = np.zeros(0, np.int32)
        self.gtt = np.zeros(0, np.int32)         # ghost tasks: target leaf, ghost leaf (batch-relative id)
        self.gts = np.zeros(0, np.int32)
        self.chunk_off = None                    # task offsets of the list groups (None: one group)
        self.timings = {}


def build_lists(pos, box, maxleaf, nside, theta=0.4, periodic=True, nthreads=0, domain_box=None, direct_start=0, nchunks=0):
    """Single-rank producer: local tree + local list + (optionally) the 26 periodic-image ghost lists.
    nchunks > 0: the local list is produced by the chunked walk plan (same task multiset, grouped by target
    chunk) and `chunk_off` gives the task offsets of the groups, for the chunk-pipelined device step."""
    rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
    out = HostLists()
    t0 = time.perf_counter()
    bdl, bdr = ([0.0] * 3, [box] * 3) if domain_box is None else domain_box
    T = host.LocalTree(pos, maxleaf, bdl, bdr, direct_start, nthreads)
    t1 = time.perf_counter()
    out.tree = T
    if periodic:
        check_wrap_condition(T, box, rcut, "build_lists")
    if nchunks > 0:
        plan = T.walk_plan(theta, rcut, nchunks)
        parts = [plan.run(c, nthreads) for c in range(plan.nchunks)]
        out.tt = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.int32)
        out.ts = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
        out.chunk_off = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])]).astype(np.int64)
    else:
        out.tt, out.ts = T.walk_task_p2p(theta, rcut, nthreads)
        out.chunk_off = np.array([0, len(out.tt)], np.int64)
    t2 = time.perf_counter()
    if periodic:
        gp, gs, gc, gtt, gts = [], [], [], [], []
        nbody = nleaf = 0
        for sh in SHIFTS:
            disp = np.array(sh, np.float64) * box
            img = T.prepare_sendtree(T.node_center[0], T.node_width[0], disp, theta, rcut)
            tt, ts = T.walk_task_p2p_ext(img, theta, rcut, nthreads)
            if len(tt) == 0:
                continue
            # compact: only image leaves that are actually referenced become ghost leaves
            used, inv = np.unique(ts, return_inverse=True)
            cnt = img.npart[used]
            st = img.son[used, 0]
            sel = np.concatenate([np.arange(s, s + c) for s, c in zip(st, cnt)]) if len(used) else np.zeros(0, np.int64)
            gp.append(img.body[sel])
            gs.append(nbody + np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32))
            gc.append(cnt.astype(np.int32))
            gtt.append(tt)
            gts.append((nleaf + inv).astype(np.int32))
            nbody += int(cnt.sum())
            nleaf += len(used)
        if gp:
            out.ghost_pos = np.concatenate(gp)
            out.ghost_start = np.concatenate(gs).astype(np.int32)
            out.ghost_count = np.concatenate(gc).astype(np.int32)
            out.gtt = np.concatenate(gtt).astype(np.int32)
            out.gts = np.concatenate(gts).astype(np.int32)
    t3 = time.perf_counter()
    out.timings = dict(build_s=t1 - t0, walk_s=t2 - t1, images_s=t3 - t2)
    out.params = dict(rs=rs, rcut=rcut, eps=eps, box=box)
    return out


def chunked_task_arrays(lists):
    """(tt, ts, chunk_off) for P2PContext.step_host_chunked: the local groups followed by ONE group holding
    the ghost tasks, whose source ids are made absolute (ghost leaf g -> nleaf + g)."""
    T = lists.tree
    off0 = lists.chunk_off if lists.chunk_off is not None else np.array([0, len(lists.tt)], np.int64)
    if len(lists.gtt) == 0:
        return lists.tt, lists.ts, off0
    tt = np.concatenate([lists.tt, lists.gtt]).astype(np.int32)
    ts = np.concatenate([lists.ts, lists.gts + T.nleaf]).astype(np.int32)
    off = np.concatenate([off0, [len(tt)]]).astype(np.int64)
    return tt, ts, off


class ShortRangeStep:
    """Device side of the step for one rank."""

    def __init__(self, device=0, variant=0):
        self.ctx = P2PContext(device)
        self.ctx.set_kernel_variant(variant)

    def upload(self, lists, mass, truncated=True):
        T, c = lists.tree, self.ctx
        prm = lists.params
        c.set_physics(mass, prm["eps"], prm["rs"] if truncated else 0.0)
        c.set_box([0.0, 0.0, 0.0], prm["box"])       # periodic frame: displaced ghosts wrap onto their originals
        c.upload_particles(T.pos)
        c.upload_leaves(T.leaf_npart, T.leaf_ipart)
        c.clear_tasks()
        c.append_tasks(lists.tt, lists.ts)
        if len(lists.gtt):
            first = c.append_ghosts(lists.ghost_pos, lists.ghost_start, lists.ghost_count)
            c.append_tasks(lists.gtt, lists.gts, source_offset=first)
        c.build_csr()

    def compute(self):
        self.ctx.compute()

    def download(self, lists):
        """Accelerations in the ORIGINAL particle order."""
        a = self.ctx.download_acc()
        out = np.empty_like(a)
        out[lists.tree.perm] = a
        return out

    def run(self, lists, mass, truncated=True):
        self.upload(lists, mass, truncated)
        self.compute()
        return self.download(lists)


def run_full_step(ctx, pos, box, maxleaf, nside, mass, theta=0.4, nchunks=16, periodic=True, truncated=True, nthreads=0,
                  pipelined=True, acc_out=None):
    """The whole short-range step -- tree build, walk, P2P.

    pipelined=True: walk/compute pipeline -- the device packs and computes target chunk c while the host
    walks chunk c+1 into a pinned double buffer (the reference's ping-pong task buffers,
    1_Indexing/src/fmm.c:365-400,947-1024, which this fork serialised).  pipelined=False: the same work
    strictly in sequence (walk everything, then upload and compute everything), for comparison.
    Periodic-image ghost tasks form one last chunk.
    Returns (acc in TREE order, tree, timings dict, ntask, npairs); acc[i] belongs to input particle tree.perm[i]."""
    import torch
    t = {}
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
    T = host.LocalTree(pos, maxleaf, [0.0] * 3, [box] * 3, 0, nthreads)
    if periodic:
        check_wrap_condition(T, box, rcut, "run_full_step")
    t["build_s"] = time.perf_counter() - t0
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    ctx.upload_particles(T.pos)
    ctx.upload_leaves(T.leaf_npart, T.leaf_ipart)
    t1 = time.perf_counter()
    t["upload_s"] = t1 - t0 - t["build_s"]
    plan = T.walk_plan(theta, rcut, nchunks)
    cap = max(1024, int(1.3 * 120 * T.nleaf / max(plan.nchunks, 1)))
    nslot = 2 if pipelined else plan.nchunks
    bufs, views = [], []
    for _ in range(nslot):
        pair = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        bufs.append(pair)
        views.append([b.numpy() for b in pair])
    done = [torch.cuda.Event() for _ in range(nslot)]
    used = [False] * nslot
    t["plan_s"] = time.perf_counter() - t1

    def walk_into(c, slot):
        if used[slot]:
            done[slot].synchronize()                  # the H2D that read this buffer has finished
        n = plan.run_into(c, views[slot][0], views[slot][1], nthreads)
        if n < 0:                                     # buffer too small for this chunk: replace it
            pair = [torch.empty(int(-n * 1.2), dtype=torch.int32).pin_memory() for _ in range(2)]
            bufs[slot] = pair
            views[slot] = [b.numpy() for b in pair]
            n = plan.run_into(c, views[slot][0], views[slot][1], nthreads)
        return n

    def submit(slot, n):
        ctx.clear_tasks()
        ctx.append_tasks(views[slot][0][:n], views[slot][1][:n])      # async H2D from pinned memory
        if ctx.stream_ptr:
            done[slot].record(torch.cuda.ExternalStream(ctx.stream_ptr))
        else:
            ctx.synchronize()
        used[slot] = True
        ctx.build_csr()
        ctx.compute()                                 # returns at once

    t2 = time.perf_counter()
    if pipelined:
        n_next = walk_into(0, 0) if plan.nchunks else 0
        for c in range(plan.nchunks):
            submit(c & 1, n_next)                     # the device works on chunk c ...
            if c + 1 < plan.nchunks:
                n_next = walk_into(c + 1, (c + 1) & 1)   # ... while the host walks chunk c+1
    else:
        counts = [walk_into(c, c) for c in range(plan.nchunks)]
        ctx.synchronize()
        for c in range(plan.nchunks):
            submit(c, counts[c])
        ctx.synchronize()
    t["walk_and_submit_s"] = time.perf_counter() - t2
    if periodic:
        t3 = time.perf_counter()
        first = None
        gtt, gts = [], []
        for sh in SHIFTS:
            img = T.prepare_sendtree(T.node_center[0], T.node_width[0], np.array(sh, np.float64) * box, theta, rcut)
            tt, ts = T.walk_task_p2p_ext(img, theta, rcut, nthreads)
            if len(tt) == 0:
                continue
            used_l, inv = np.unique(ts, return_inverse=True)
            cnt, st = img.npart[used_l], img.son[used_l, 0]
            sel = np.concatenate([np.arange(s, s + k) for s, k in zip(st, cnt)])
            f = ctx.append_ghosts(img.body[sel], np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32), cnt)
            first = f if first is None else first
            gtt.append(tt)
            gts.append((f - first + inv).astype(np.int32))
        if gtt:
            ctx.clear_tasks()
            ctx.append_tasks(np.concatenate(gtt), np.concatenate(gts), source_offset=first)
            ctx.build_csr()
            ctx.compute()
        t["images_s"] = time.perf_counter() - t3
    t4 = time.perf_counter()
    ntask, npairs = ctx.accumulated_counts()
    acc = ctx.download_acc(acc_out)
    t["download_s"] = time.perf_counter() - t4
    t["total_s"] = time.perf_counter() - t0
    return acc, T, t, ntask, npairs


def run_device_step(ctx, pos, box, maxleaf, nside, mass, theta=0.4, periodic=True, truncated=True, acc_out=None):
    """The whole short-range step with the list producers ON THE DEVICE (csrc/device_tree.cuh): one upload of
    the positions, tree build, dual-tree walk incl. the 26 periodic images, packing, forces, one download.
    Same tree, same task multiset and the same forces (up to FP32 summation order) as run_full_step.
    Returns (acc in the ORDER OF `pos`, timings dict, ntask, npairs)."""
    t0 = time.perf_counter()
    rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
    ctx.set_physics(mass, eps, rs if truncated else 0.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    acc = ctx.step_device(pos, maxleaf, [0.0] * 3, [box] * 3, theta, rcut, box if periodic else 0.0, 0, acc_out)
    total = time.perf_counter() - t0
    info = ctx.tree_info()
    st = ctx.step_timings()                      # summed over the target chunks of the step
    ntask, npairs = ctx.accumulated_counts()
    t = dict(total_s=total, build_ms=st["build_ms"], walk_ms=st["walk_ms"], csr_ms=st["csr_ms"], force_ms=st["force_ms"], chunks=st["chunks"],
             tree_levels=info["nlevel"], leaves=info["nleaf"])
    return acc, t, ntask, npairs
