#!/usr/bin/env python
"""bench.py -- P2P pair-interactions/s of the short-range step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the CPU arm (fp64 oracle port, all host threads)

Workload: synthetic LambdaCDM-like box (grid + Gaussian displacement, seed 20250101), 256^3
particles by default (BASELINE configs[1]), MAXLEAF 32, theta 0.4, erfc-truncated kernel, the 26
periodic images included.  N = 1: the whole box on one GPU.  N > 1: the SAME box split by the
reference's rank kd-tree (strong scaling), halos exchanged once per step setup over NCCL; every rank
computes the forces of the particles it owns, no data-path collective.

A step = one pass of the hot path over the whole list: `value` times the force kernel with all inputs
resident in HBM; `e2e` times the reference-facing C-ABI sequence with HOST (pinned) buffers: H2D of
positions / leaves / tasks / ghosts, device CSR packing, force kernel, D2H of accelerations.
The lists (tree build, dual-tree walk, halo images) are produced on the host by libp2p_host.so
BEFORE the timed region; their wall times are reported in config.host_setup_s.
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


def make_workload(nside, clustered=False):
    from p2p_b200 import synth
    if clustered:
        pos, box = synth.clustered(nside)
    else:
        pos, box = synth.zeldovich_like(nside)
    return pos, box, synth.DEMO_MASS


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
    """fp64 oracle (the CPU port of the reference's pair arithmetic) on a bounded sample of rows."""
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
    return (n1 + n2) / dt, n1 + n2, nrow, dt, oracle.max_threads() if nthreads <= 0 else nthreads


def run_reference_arm(args):
    """--impl reference: the reference has no CPU P2P source, so the arm is the oracle
    port of its pair arithmetic over the reference's lists, OpenMP over all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from p2p_b200 import step
    pos, box, mass = make_workload(args.nside, args.clustered)
    L = step.build_lists(pos, box, args.maxleaf, args.nside, THETA, periodic=True, nthreads=os.cpu_count() or 1)
    rates, tot_pairs, tot_dt = [], 0, 0.0
    # one step = a bounded sample (default 5e9 pairs ~ 9 s on 16 cores) so that W + K steps end within minutes
    per_step_pairs = min(args.cpu_pairs, 5e9)
    for i in range(args.warmup + args.steps):
        rate, npairs, nrow, dt, threads = cpu_oracle_rate(L, mass, per_step_pairs, os.cpu_count() or 1)
        if i >= args.warmup:
            rates.append(rate); tot_pairs += npairs; tot_dt += dt
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


def workload_config(args, box):
    return {"workload": f"{args.nside}^3 particles {'clustered (Zeldovich + NFW-like clumps)' if args.clustered else 'Zeldovich-like (grid + sigma 0.2 spacing)'}"
                        f", box {box:g} h^-1 kpc, MAXLEAF {args.maxleaf}, theta {THETA}, NSIDE {args.nside}, erfc-truncated kernel, "
                        f"local list + 26 periodic images", "nside": args.nside, "maxleaf": args.maxleaf, "seed": 20250101,
            "l2_policy": "inputs larger than L2 (particles + CSR columns >> 126 MB); no flush between steps"}


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
    ap.add_argument("--chunks", type=int, default=16, help="target chunks of the list (groups of the pipelined host step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-step", action="store_true", help="skip the walk/compute pipeline measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from p2p_b200 import step
    from p2p_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the P2P library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---------------------------------------------------------------- host setup (untimed)
    t0 = time.perf_counter()
    pos, box, mass = make_workload(args.nside, args.clustered)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    # torchrun exports OMP_NUM_THREADS=1; the host producers take an explicit thread count instead
    nthreads = max(1, (os.cpu_count() or 1) // world)
    if distributed:
        L = pdist.build_lists(pos, box, args.maxleaf, args.nside, THETA, nthreads=nthreads)
    else:
        L = step.build_lists(pos, box, args.maxleaf, args.nside, THETA, periodic=True, nthreads=nthreads, nchunks=args.chunks)
    t_lists = time.perf_counter() - t0
    T = L.tree

    st = step.ShortRangeStep(local_rank)
    ctx = st.ctx
    # a dedicated torch stream (the legacy default stream has handle 0, which the C-ABI reads as "own stream"):
    # the kernels are launched on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # pinned host buffers for the e2e leg
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    hp = {}
    for name, arr in (("pos", T.pos), ("leaf_npart", T.leaf_npart), ("leaf_ipart", T.leaf_ipart), ("tt", L.tt), ("ts", L.ts),
                      ("gpos", L.ghost_pos.reshape(-1, 3)), ("gstart", L.ghost_start), ("gcount", L.ghost_count), ("gtt", L.gtt),
                      ("gts", L.gts)):
        t, v = pin(arr)
        keep.append(t)
        hp[name] = v
    acc_t = torch.empty((T.npart, 3), dtype=torch.float64).pin_memory()
    acc_host = acc_t.numpy()
    h2d = sum(hp[k].nbytes for k in hp)
    d2h = acc_host.nbytes

    ctt, cts, coff = step.chunked_task_arrays(L)
    t1_, ctt = pin(ctt)
    t2_, cts = pin(cts)
    keep += [t1_, t2_]
    h2d = sum(hp[k].nbytes for k in ("pos", "leaf_npart", "leaf_ipart", "gpos", "gstart", "gcount")) + ctt.nbytes + cts.nbytes

    def resident_setup():
        """everything on the device, ONE CSR over the whole list: the state the kernel-only leg times"""
        ctx.set_physics(mass, L.params["eps"], L.params["rs"])
        ctx.set_box([0.0, 0.0, 0.0], box)
        ctx.upload_particles(hp["pos"])
        ctx.upload_leaves(hp["leaf_npart"], hp["leaf_ipart"])
        ctx.clear_tasks()
        ctx.append_tasks(hp["tt"], hp["ts"])
        if len(hp["gtt"]):
            first = ctx.append_ghosts(hp["gpos"], hp["gstart"], hp["gcount"])
            ctx.append_tasks(hp["gtt"], hp["gts"], source_offset=first)
        ctx.build_csr()

    def e2e_step():
        """the reference-facing call: host (pinned) buffers in, host accelerations out; H2D + packing of list
        group g+1 overlap the kernel of group g (p2p_step_host_chunked)"""
        ctx.set_physics(mass, L.params["eps"], L.params["rs"])
        ctx.set_box([0.0, 0.0, 0.0], box)
        ctx.step_host_chunked(hp["pos"], hp["leaf_npart"], hp["leaf_ipart"], ctt, cts, coff, hp["gpos"], hp["gstart"],
                              hp["gcount"], acc=acc_host)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # resident state for the kernel-only leg
    resident_setup()
    ntask, npairs = ctx.counts()

    # ---------------------------------------------------------------- value: kernel with resident inputs
    for _ in range(max(args.warmup, 3)):
        ctx.zero_acc(); ctx.compute()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        ctx.zero_acc(); ctx.compute()
        ev[i + 1].record(stream)
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    # launch duration of the force kernel alone (library events around the launch, same stream)
    kernel_ms = ctx.last_timings()[0]

    # ---------------------------------------------------------------- e2e: host buffers through the C-ABI
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    resident_setup()
    ctx.synchronize()
    _, csr_ms = ctx.last_timings()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---------------------------------------------------------------- full step with the walk/compute pipeline
    full_step = None
    if not distributed and not args.no_full_step:
        ctx2 = step.ShortRangeStep(local_rank).ctx
        ctx2.set_stream(stream.cuda_stream)
        acc2 = torch.empty((T.npart, 3), dtype=torch.float64).pin_memory().numpy()
        res = {}
        for mode in (False, True, False, True):        # sequential / pipelined, twice: the second pair is reported
            barrier()
            t0 = time.perf_counter()
            _, _, tp, n_t, n_p = step.run_full_step(ctx2, pos, box, args.maxleaf, args.nside, mass, THETA, nchunks=32,
                                                    periodic=True, nthreads=nthreads, pipelined=mode, acc_out=acc2)
            barrier()
            res[mode] = (time.perf_counter() - t0, tp)
            assert (n_t, n_p) == (ntask, npairs), ((n_t, n_p), (ntask, npairs))
        # the same step with the list producers on the device (tree build + dual-tree walk kernels)
        ppos = torch.from_numpy(pos).pin_memory().numpy()
        dev = None
        for _ in range(4):
            barrier()
            t0 = time.perf_counter()
            _, td, n_t, n_p = step.run_device_step(ctx2, ppos, box, args.maxleaf, args.nside, mass, THETA, periodic=True, acc_out=acc2)
            barrier()
            dev = (time.perf_counter() - t0, td)
            assert (n_t, n_p) == (ntask, npairs), ((n_t, n_p), (ntask, npairs))
        del ppos
        full_step = {"what": "positions in, accelerations out: tree build + dual-tree walk + 26 periodic-image walks + CSR + P2P, one rank",
                     "device_resident_s": dev[0], "device_resident_breakdown": dev[1],
                     "host_pipelined_s": res[True][0], "host_pipelined_breakdown": res[True][1],
                     "host_sequential_s": res[False][0], "host_sequential_breakdown": res[False][1], "host_threads": nthreads,
                     "note": "device_resident: list producers as CUDA kernels (p2p_step_device); host_*: list producers on the host "
                             "cores (libp2p_host.so), pipelined = host walks chunk c+1 while the device computes chunk c"}
        ctx2.close()
    elif not args.no_full_step:
        # multi-rank: every rank builds its tree and walks it against the trees of all ranks on its GPU; halo particles
        # travel leaf-granular over NCCL (p2p_b200/dist_device.py)
        from p2p_b200 import dist_device
        lp, _, tcenter, twidth, direct, dom = pdist.decompose(pos, box)
        c_, w_ = tcenter[dom], twidth[dom]
        lp_t, lp_pin = pin(lp)
        ctx2 = step.ShortRangeStep(local_rank).ctx
        ctx2.set_stream(stream.cuda_stream)
        acc2 = torch.empty((lp.shape[0], 3), dtype=torch.float64).pin_memory().numpy()
        dev = None
        for _ in range(4):
            tm = {}
            barrier()
            t0 = time.perf_counter()
            _, n_t, n_p = dist_device.run_device_step(ctx2, lp_pin, pos.shape[0], box, args.maxleaf, args.nside, mass, c_ - 0.5 * w_,
                                                      c_ + 0.5 * w_, int(direct[dom]), THETA, periodic=True, truncated=True,
                                                      acc_out=acc2, timings=tm)
            barrier()
            dev = (time.perf_counter() - t0, tm)
            assert (n_t, n_p) == (ntask, npairs), ((n_t, n_p), (ntask, npairs))
        full_step = {"what": "per-rank positions in, accelerations out: tree build + topology all-gather + walk against every rank's tree "
                             "(27 displacements) + leaf-granular halo fetch (NCCL all-to-all-v) + CSR + P2P; wall time between barriers",
                     "device_resident_s": dev[0], "rank0_breakdown": dev[1]}
        ctx2.close()
    del pos

    # ---------------------------------------------------------------- reduce over ranks
    if distributed:
        t = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kernel_ms_max = t.tolist()
        s = torch.tensor([npairs, ntask, h2d, d2h, T.npart], dtype=torch.float64, device="cuda")
        mx = s.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        all_pairs, all_tasks, all_h2d, all_d2h, all_part = [int(x) for x in s.tolist()]
        imbalance = 1.0 - all_pairs / (world * mx[0].item())       # the reference's definition, 1_Indexing/src/photoNs.c:309
    else:
        all_pairs, all_tasks, all_h2d, all_d2h, all_part, imbalance, kernel_ms_max = npairs, ntask, h2d, d2h, T.npart, 0.0, kernel_ms

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        props = torch.cuda.get_device_properties(local_rank)
        fp32_peak = props.multi_processor_count * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12     # TFLOP/s
        value = all_pairs * args.steps / (total_ms * 1e-3)
        # roofline of the dominant kernel: this rank's pairs over its own launch duration
        achieved = npairs * FLOP_PER_PAIR / (kernel_ms * 1e-3) / 1e12
        algo_bytes = T.npart * 16 + T.npart * 16 + ntask * 4 + (len(L.ghost_pos)) * 16
        out = {
            "metric": "P2P pair-interactions/s", "value": value, "unit": "pair/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, box), parallelism=f"domains{args.gpus}", particles=all_part, tasks=all_tasks,
                           pairs_per_step=all_pairs, imbalance=imbalance,
                           host_setup_s={"generate": t_gen, "lists_total": t_lists, **L.timings} if not distributed
                           else {"generate": t_gen, "lists_total": t_lists}),
            "e2e": {"value": all_pairs * args.steps / e2e_s, "unit": "pair/s", "h2d_bytes_per_step": all_h2d,
                    "d2h_bytes_per_step": all_d2h, "ms_per_step": 1e3 * e2e_s / args.steps, "csr_pack_ms": csr_ms},
            "gpu_launches": 2 * args.steps,       # per timed step: p2p_rows_kernel + add_counter_kernel (the pair count), on this rank
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": 1.35e9 if (not distributed and args.nside == 256 and args.maxleaf == 32 and not args.clustered) else None,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch at this workload (profiles/r1e_ncu_rows_kernel_256_final.txt); null for other workloads", "kernel": "p2p_rows_kernel<TT=16,NSRC=2,STAGE=384,trunc,packed(FFMA2),4 blocks/SM,split polynomial>", "kernel_ms": kernel_ms,
                         "flop_per_pair": FLOP_PER_PAIR,
                         "peak_source": f"{props.multi_processor_count} SMs x 128 FP32 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json ({peaks_kind}); "
                                        "the file holds no FP32 figure, SURVEY.md section 8d defines this peak",
                         "hbm": {"algorithmic_bytes": algo_bytes, "achieved_gbs": algo_bytes / (kernel_ms * 1e-3) / 1e9,
                                 "peak_gbs": peaks["hbm_gbs"], "note": "compulsory particle/CSR/acc traffic of one launch; the kernel is FP32-pipe bound"}},
            "clocks": sampler.result(),
        }
        if full_step:
            out["config"]["full_step"] = full_step
        if not args.no_cpu_baseline and not distributed:
            rate, npr, nrow, dt, threads = cpu_oracle_rate(L, mass, args.cpu_pairs, os.cpu_count() or 1)
            rate1, npr1, _, dt1, _ = cpu_oracle_rate(L, mass, max(args.cpu_pairs / 40, 1e8), 1)
            out["cpu_baseline"] = {"value": rate, "unit": "pair/s", "cores": threads, "kind": "port",
                                   "sample": f"first {nrow} of {T.nleaf} target leaves (complete CSR rows) = {npr} pairs in {dt:.1f} s, "
                                             "fp64 oracle (erfc/exp), OpenMP dynamic over target leaves",
                                   "single_thread": {"value": rate1, "unit": "pair/s", "sample": f"{npr1} pairs in {dt1:.1f} s"}}
        print_json(out)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
