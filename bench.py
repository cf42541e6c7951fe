#!/usr/bin/env python
"""bench.py -- P2P pair-interactions/s of the short-range step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the CPU arm (fp64 oracle port, all host threads)

Workload: synthetic LambdaCDM-like box (grid + Gaussian displacement, seed 20250101), 256^3 particles by default
(BASELINE configs[1]), MAXLEAF 32, theta 0.4, erfc-truncated kernel, the 26 periodic images included.  N = 1: the whole
box on one GPU.  N > 1: the SAME box split by the reference's rank kd-tree (strong scaling); every rank generates only the
slab of the global particle array it starts from, the slabs are routed on the devices, and every rank computes the forces
of the particles it owns.

A step = one short-range step of the hot path, EVERYTHING per-step inside the timed region at every N:
  value : particles resident in HBM (in the order the previous step left them) -> tree build, dual-tree walk (own tree +
          26 images), [N > 1: topology all-gather, walk against the other ranks' trees, leaf-granular halo exchange over
          NCCL], list packing, force kernels.  Device time (CUDA events), max over ranks.
  e2e   : the same step through the reference-facing call with HOST buffers: pinned host positions in (H2D), pinned host
          accelerations out (D2H), wall clock between barriers, max over ranks.
  roofline : the force kernel alone (library events around its launches), 38 flop per pair against the FP32 FMA peak.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"),):
    if p not in sys.path:
        sys.path.insert(0, p)

FLOP_PER_PAIR = 38.0      # algorithmic FP32 flop of one truncated pair (FMA = 2)
THETA = 0.4
print_json = None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        # samples under load: the upper half (the sampler also sees the idle gaps between launches)
        med = s[len(s) * 3 // 4] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def generate(args, lo, hi):
    from p2p_b200 import synth
    gen = synth.clustered_slab if args.clustered else synth.zeldovich_slab
    return gen(args.nside, lo, hi)


def workload_config(args, box):
    return {"workload": f"{args.nside}^3 particles {'clustered (Zeldovich + NFW-like clumps)' if args.clustered else 'Zeldovich-like (grid + sigma 0.2 spacing)'}"
                        f", box {box:g} h^-1 kpc, MAXLEAF {args.maxleaf}, theta {THETA}, NSIDE {args.nside}, erfc-truncated kernel, "
                        f"local list + 26 periodic images", "nside": args.nside, "maxleaf": args.maxleaf, "seed": 20250101,
            "l2_policy": "inputs larger than L2 (particles + CSR columns >> 126 MB); no flush between steps"}


def host_lists(args, pos, box, nthreads):
    from p2p_b200 import step
    return step.build_lists(pos, box, args.maxleaf, args.nside, THETA, periodic=True, nthreads=nthreads)


def sample_rows_for_cpu(L, target_pairs):
    """Leading target leaves of the list holding ~target_pairs pairs (local + ghost tasks), as arrays
    for the oracle: complete CSR rows, so the CPU does exactly the work the GPU does for them."""
    T = L.tree
    per_task = T.leaf_npart[L.tt].astype(np.int64) * T.leaf_npart[L.ts]
    per_row = np.bincount(L.tt, weights=per_task, minlength=T.nleaf)
    if len(L.gtt):
        per_row += np.bincount(L.gtt, weights=T.leaf_npart[L.gtt].astype(np.int64) * L.ghost_count[L.gts], minlength=T.nleaf)
    nrow = int(np.searchsorted(np.cumsum(per_row), target_pairs)) + 1
    nrow = min(nrow, T.nleaf)
    m = L.tt < nrow
    mg = L.gtt < nrow
    return nrow, (L.tt[m], L.ts[m]), (L.gtt[mg], L.gts[mg])


def cpu_oracle_rate(L, mass, target_pairs, nthreads):
    """fp64 oracle (the CPU port of the reference's pair arithmetic) on a bounded sample of rows.
    Returns (pair/s, pairs, rows, seconds, threads, accelerations [tree order; only the sampled rows are filled])."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    T = L.tree
    nrow, (tt, ts), (gtt, gts) = sample_rows_for_cpu(L, target_pairs)
    acc = np.zeros((T.npart, 3))
    t0 = time.perf_counter()
    _, n1 = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass, L.params["eps"],
                       L.params["rs"], acc=acc, nthreads=nthreads)
    n2 = 0
    if len(gtt):
        _, n2 = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, L.ghost_pos, L.ghost_count, L.ghost_start, gtt, gts, mass,
                           L.params["eps"], L.params["rs"], acc=acc, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return (n1 + n2) / dt, n1 + n2, nrow, dt, oracle.max_threads() if nthreads <= 0 else nthreads, acc


def run_reference_arm(args):
    """--impl reference: the reference has no CPU P2P source, so the arm is the oracle
    port of its pair arithmetic over the reference's lists, OpenMP over all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from p2p_b200 import synth
    pos, box = generate(args, 0, args.nside ** 3)
    mass = synth.DEMO_MASS
    L = host_lists(args, pos, box, os.cpu_count() or 1)
    tot_pairs, tot_dt = 0, 0.0
    # one step = a bounded sample (default 5e9 pairs ~ 9 s on 16 cores) so that W + K steps end within minutes
    per_step_pairs = min(args.cpu_pairs, 5e9)
    for i in range(args.warmup + args.steps):
        rate, npairs, nrow, dt, threads, _ = cpu_oracle_rate(L, mass, per_step_pairs, os.cpu_count() or 1)
        if i >= args.warmup:
            tot_pairs += npairs; tot_dt += dt
    value = tot_pairs / tot_dt
    out = {
        "impl": "reference", "metric": "P2P pair-interactions/s", "value": value, "unit": "pair/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, box),
        "cpu_baseline": {"value": value, "unit": "pair/s", "cores": threads, "kind": "port",
                         "sample": f"first {nrow} target leaves (complete CSR rows, local + periodic-image tasks) = "
                                   f"{tot_pairs // args.steps} pairs per step, fp64 erfc/exp, OpenMP dynamic over target leaves"},
        "e2e": {"value": value, "unit": "pair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print_json(out)


def count_kernels(fn):
    """kernels of this process launched by one call of fn (CUPTI through torch.profiler); NCCL's own kernels are listed apart"""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if str(e.device_type).endswith("CUDA") and not e.name.startswith(("Memcpy", "Memset"))]
        ours = [n for n in names if "nccl" not in n.lower()]
        return len(ours), len(names) - len(ours)
    except Exception as e:  # pragma: no cover
        print(f"[bench] kernel count unavailable: {type(e).__name__}: {e}", file=sys.stderr)
        return None, None


def roofline_traffic(args, world):
    """DRAM bytes of one launch of the force kernel at this workload, from the committed ncu capture (profiles/)"""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.isfile(path):
        return None, "no ncu capture committed for this workload"
    with open(path) as f:
        tab = json.load(f)
    key = f"{args.nside}{'c' if args.clustered else ''}_ml{args.maxleaf}_n{world}"
    if key in tab:
        return tab[key]["dram_bytes"], tab[key]["source"]
    return None, f"no ncu capture for {key} in profiles/roofline_traffic.json"


def main():
    # Only the JSON line may reach stdout: NCCL (version banner) and the reference-style prints of the
    # libraries go to fd 1 as well, so fd 1 is pointed at stderr and the JSON is written to the saved fd.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_json
    print_json = lambda obj: (real_stdout.write(json.dumps(obj) + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nside", type=int, default=256)
    ap.add_argument("--maxleaf", type=int, default=32)
    ap.add_argument("--clustered", action="store_true")
    ap.add_argument("--cpu-pairs", type=float, default=1.5e10, help="pairs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (sizes whose positions do not fit pinned host memory)")
    ap.add_argument("--no-overlap", action="store_true", help="remote phase on the compute stream (A/B of the halo overlap)")
    ap.add_argument("--no-launch-count", action="store_true", help="skip the CUPTI kernel count (it cannot run under ncu)")
    ap.add_argument("--relax", type=int, default=0, help="work-weighted split relaxations before the timed region (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import p2p_b200
    from p2p_b200 import dist_device, host, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the P2P library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 or "RANK" in os.environ:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)      # NCCL's kernels outrank the force kernel
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group("gloo", rank=0, world_size=1)                   # one rank: no collective is ever issued
    nvtx = torch.cuda.nvtx

    # ---------------------------------------------------------------- setup (untimed): slab, routing, resident particles
    npart = args.nside ** 3
    box = synth.box_for(args.nside)
    mass = synth.DEMO_MASS
    rs, rcut, eps = host.derived_params(box, args.nside, npart)
    lo, hi = npart * rank // world, npart * (rank + 1) // world
    split = host.domain_setup(world, box)[0]
    ctx = p2p_b200.P2PContext(local_rank)
    ctx.set_physics(mass, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], box)
    S = dist_device._streams_of(ctx)
    stream = S.main
    torch.cuda.set_stream(stream)
    overlap = not args.no_overlap
    # the slab goes to the device in pieces: neither the host nor the staging buffer ever holds more than 2^24 particles
    # (1024^3 on one GPU is 26 GB of fp64 positions); only the e2e leg at N = 1 keeps a host copy of the box
    t0 = time.perf_counter()
    keep_host = world == 1 and not args.no_e2e
    pieces = []
    for a in range(lo, hi, 1 << 24):
        b = min(hi, a + (1 << 24))
        piece, _ = generate(args, a, b)
        ctx.route_load(piece, a, append=a > lo)
        if keep_host:
            pieces.append(piece)
    slab = np.concatenate(pieces) if keep_host else None
    del pieces
    t_gen = time.perf_counter() - t0

    def domain():
        center, width, direct = host.domain_boxes(world, box, split)
        dom = host.domain_of_rank(world, rank)
        return center[dom] - 0.5 * width[dom], center[dom] + 0.5 * width[dom], int(direct[dom])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    nloc = dist_device.migrate(ctx, world, split, None, dev)
    bdl, bdr, direct = domain()
    tm = {}

    def resident_step(timings=None):
        """the timed unit of `value`: particles resident in HBM -> accelerations resident in HBM"""
        nvtx.range_push("resident_step")
        ctx.tree_build_resident(args.maxleaf, bdl, bdr, direct)
        out = dist_device.lists_halo_forces(ctx, rcut, box, bdl, bdr, THETA, True, None, timings, False, True, overlap)
        nvtx.range_pop()
        return out

    ntask, npairs = resident_step()
    history = []
    for _ in range(args.relax if world > 1 else 0):
        # the reference's feedback (1_Indexing/src/photoNs.c:295-306): splits move towards equal work, particles migrate
        w_all = dist_device._gather_host_ints([ntask], None, dev)[:, 0].astype(np.float64)
        split = host.domain_relax(world, box, split, w_all)
        nloc = dist_device.migrate(ctx, world, split, None, dev)
        bdl, bdr, direct = domain()
        ntask, npairs = resident_step()
        history.append(float(1.0 - w_all.sum() / (world * w_all.max())))
    t_setup = time.perf_counter() - t0

    # ---------------------------------------------------------------- value: device-resident step, per-step communication inside
    for _ in range(max(args.warmup, 3) - 1):
        resident_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kernel_ms = []
    for i in range(args.steps):
        tm = {}
        ntask, npairs = resident_step(tm)
        kernel_ms.append(tm["force_ms"])
    e1.record(stream)
    barrier()
    value_ms = e0.elapsed_time(e1) / args.steps
    force_ms = float(np.mean(kernel_ms))
    breakdown = {k: tm[k] for k in ("build_ms", "walk_ms", "csr_ms", "force_ms", "comm_ms", "remote_walk_ms", "force_local_ms", "force_remote_ms")}
    ghost_particles, chunks = tm["ghost_particles"], tm["chunks"]
    free_b, total_b = torch.cuda.mem_get_info(dev)
    hbm_used_gb = round((total_b - free_b) / 1e9, 2)

    # ---------------------------------------------------------------- e2e: host buffers through the public call
    e2e_ms, e2e_tm, h2d, d2h = None, {}, 0, 0
    if not args.no_e2e:
        if world > 1:
            bufs = [torch.empty(nloc, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(nloc, dtype=torch.int32, device=dev)]
            ctx.route_export(*[b.data_ptr() for b in bufs])
            hpos_t = torch.stack(bufs[:3], dim=1).cpu().pin_memory()         # this rank's particles, as a host code would hold them
            del bufs
        else:
            hpos_t = torch.from_numpy(slab).pin_memory()                     # the box in the order it was generated
            del slab
        hpos, acc_t = hpos_t.numpy(), torch.empty((nloc, 3), dtype=torch.float64).pin_memory()
        h2d, d2h = hpos.nbytes, acc_t.numpy().nbytes

        def e2e_step(timings=None):
            nvtx.range_push("e2e_step")
            out = dist_device.run_device_step(ctx, hpos, npart, box, args.maxleaf, args.nside, mass, bdl, bdr, direct, THETA, periodic=True,
                                              truncated=True, acc_out=acc_t.numpy(), timings=timings, overlap=overlap)
            nvtx.range_pop()
            return out

        for _ in range(min(max(args.warmup, 1), 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_tm = {}
            _, nt2, np2 = e2e_step(e2e_tm)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        # (N = 1: the e2e tree is built from the generation order, the resident one from the previous step's order; the split
        # means are sums in those orders, so a leaf pair at the very edge of the cut may differ)
        assert abs(np2 - npairs) <= 1e-6 * npairs, ((nt2, np2), (ntask, npairs))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches, nccl_launches = (None, None) if args.no_launch_count else count_kernels(resident_step)

    # ---------------------------------------------------------------- reduce over ranks
    if world > 1:
        t = torch.tensor([value_ms, e2e_ms or 0.0, force_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_ms, e2e_max, force_ms_max = t.tolist()
        e2e_ms = e2e_max if e2e_ms is not None else None
        s = torch.tensor([npairs, ntask, h2d, d2h, nloc, ghost_particles], dtype=torch.float64, device=dev)
        mx = s.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        all_pairs, all_tasks, all_h2d, all_d2h, all_part, all_ghost = [int(x) for x in s.tolist()]
        imbalance = 1.0 - all_pairs / (world * mx[0].item())       # the reference's definition, 1_Indexing/src/photoNs.c:309
    else:
        all_pairs, all_tasks, all_h2d, all_d2h, all_part, all_ghost, imbalance = npairs, ntask, h2d, d2h, nloc, 0, 0.0

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        props = torch.cuda.get_device_properties(local_rank)
        fp32_peak = props.multi_processor_count * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12     # TFLOP/s
        # roofline of the dominant kernel: this rank's pairs over the duration of its force-kernel launches
        achieved = npairs * FLOP_PER_PAIR / (force_ms * 1e-3) / 1e12
        algo_bytes = nloc * 32 + ntask * 4 + ghost_particles * 16
        traffic, traffic_source = roofline_traffic(args, world)
        out = {
            "metric": "P2P pair-interactions/s", "value": all_pairs / (value_ms * 1e-3), "unit": "pair/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": value_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, box), parallelism=f"domains{args.gpus}", particles=all_part, tasks=all_tasks,
                           pairs_per_step=all_pairs, imbalance=imbalance, ghost_particles=all_ghost, chunks=chunks,
                           halo_overlap=bool(overlap and world > 1), relaxations=len(history),
                           step="resident particles -> tree build + walk + [topology / halo exchange] + packing + forces",
                           setup_s=round(t_setup, 3), generate_and_upload_s=round(t_gen, 3), hbm_used_gb_rank0=hbm_used_gb),
            "step_breakdown_ms": breakdown,
            "gpu_launches": (launches * args.steps) if launches is not None else None,
            "gpu_launches_per_step": launches, "nccl_kernels_per_step": nccl_launches,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": traffic, "traffic_source": traffic_source,
                         "kernel": "p2p_rows2_kernel<NSRC=1,STAGE=384,trunc,3 blocks/SM> (one pass per row, near + far bodies, FFMA2)",
                         "kernel_ms": force_ms, "flop_per_pair": FLOP_PER_PAIR,
                         "note": "kernel_ms = sum of the force-kernel launches of one step inside the timed region; with the halo overlap "
                                 "(N > 1) the local kernel shares the SMs with the walk / packing / NCCL kernels of the comm stream",
                         "peak_source": f"{props.multi_processor_count} SMs x 128 FP32 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json ({peaks_kind}); "
                                        "the file holds no FP32 figure, SURVEY.md section 8d defines this peak",
                         "hbm": {"algorithmic_bytes": algo_bytes, "achieved_gbs": algo_bytes / (force_ms * 1e-3) / 1e9,
                                 "peak_gbs": peaks["hbm_gbs"], "note": "compulsory particle/CSR/acc traffic of one step's launches; the kernel is FP32-pipe bound"}},
            "clocks": sampler.result(),
        }
        if e2e_ms is not None:
            out["e2e"] = {"value": all_pairs / (e2e_ms * 1e-3), "unit": "pair/s", "h2d_bytes_per_step": all_h2d, "d2h_bytes_per_step": all_d2h,
                          "ms_per_step": e2e_ms, "build_ms": e2e_tm.get("build_ms"), "walk_ms": e2e_tm.get("walk_ms"),
                          "csr_ms": e2e_tm.get("csr_ms"), "force_ms": e2e_tm.get("force_ms"), "comm_ms": e2e_tm.get("comm_ms"),
                          "what": "pinned host positions of the rank's particles in, pinned host accelerations out (dist_device.run_device_step)"}
        if history:
            out["config"]["imbalance_tasks_history"] = history
        if not args.no_cpu_baseline and world == 1 and args.nside <= 256 and not args.no_e2e:
            # the CPU port on the leading rows of the same list; its accelerations double as a parity sample of THIS run
            pos, _ = generate(args, 0, npart)
            L = host_lists(args, pos, box, os.cpu_count() or 1)
            rate, npr, nrow, dt, threads, ref = cpu_oracle_rate(L, mass, args.cpu_pairs, os.cpu_count() or 1)
            rate1, npr1, _, dt1, _, _ = cpu_oracle_rate(L, mass, max(args.cpu_pairs / 40, 1e8), 1)
            out["cpu_baseline"] = {"value": rate, "unit": "pair/s", "cores": threads, "kind": "port",
                                   "sample": f"first {nrow} of {L.tree.nleaf} target leaves (complete CSR rows) = {npr} pairs in {dt:.1f} s, "
                                             "fp64 oracle (erfc/exp), OpenMP dynamic over target leaves",
                                   "single_thread": {"value": rate1, "unit": "pair/s", "sample": f"{npr1} pairs in {dt1:.1f} s"}}
            T = L.tree
            nsel = int(T.leaf_ipart[nrow - 1] + T.leaf_npart[nrow - 1])
            # the e2e leg left the accelerations in the order of hpos (= pos); the host tree built from the same array is the
            # device tree bit for bit, so its permutation maps them to the oracle's rows
            got = acc_t.numpy()[T.perm[:nsel]]
            d = np.linalg.norm(got - ref[:nsel], axis=1)
            na = np.linalg.norm(ref[:nsel], axis=1)
            out["parity"] = {"particles": nsel, "pairs": npr, "tasks_device": int(nt2), "tasks_host_lists": int(len(L.tt) + len(L.gtt)),
                             "e1_max_err_over_max_of_own_and_mean_force": float((d / np.maximum(na, na.mean())).max()),
                             "median_err_over_mean_force": float(np.median(d) / na.mean()), "tolerance": 1e-5,
                             "against": "fp64 oracle on the identical list (the cpu_baseline sample)"}
        print_json(out)
    if world > 1:
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
