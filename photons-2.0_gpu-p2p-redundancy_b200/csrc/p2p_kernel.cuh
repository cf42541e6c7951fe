// p2p_kernel.cuh -- the sm_100a near-field force kernel.
//
// Replaces ComputeP2PIndexing (1_Indexing/src/photoNs_CUDA.cu:250-387, one FP64 thread per
// (target leaf, source leaf) task, sources in a per-thread stack array) and ComputeP2PDualNaive /
// ComputeP2PSelfInteractions (2_Redundant/src/photoNs_CUDA.cu:225-309, 386-458).
//
// Data: particles live in HBM as int4 {xi, yi, zi, mass bits}: 32-bit FIXED-POINT coordinates over
// the (padded) box, so resolution is box/2^32 whatever the box size (FP32 absolute coordinates lose
// ~6e-8*box, SURVEY hard part H2) and periodic images are free: the difference of two fixed-point
// coordinates wraps to the minimal image, so image / halo sources need no displaced copies.
//
// Work decomposition (one warp = one CSR row = one target leaf at a time, persistent warps pull rows
// from an atomic counter):
//   * the row's source leaves are streamed through a private double-buffered shared-memory ring:
//     each lane owns one source leaf of the chunk and issues ONE bulk async copy (cp.async.bulk,
//     SASS UBLKCP) of that leaf's particle run; completion is tracked by an mbarrier transaction
//     count, so a warp never blocks on a load it issued itself;
//   * SOURCES are spread over the lanes (NSRC per lane), TARGETS are walked by an unrolled loop with
//     the target coordinates broadcast from shared memory and the per-target accumulators held in
//     registers.  Lane utilisation is therefore independent of the leaf occupancy (a lane = target
//     mapping idles 14-30 % of the lanes at the reference's leaf fill of 57-76 %),
//     and the 32-lane reduction happens once per row, not once per tile;
//   * when a source is loaded from the ring it is converted to FP32 RELATIVE to the row's first
//     target (integer subtract, I2F, scale): all FP32 arithmetic is on short separations;
//   * partial slices are carried across chunk boundaries in registers, so only the last slice of a
//     row is ragged (padded with a dummy source whose contribution is exactly 0).
// Pair arithmetic (2_Redundant/src/photoNs_CUDA.cu:432-450 with the eps branch of
// 1_Indexing/src/photoNs_CUDA.cu:346-350 as an fmax on r^2), in units of 2 r_s / sqrt(log2 e) so that
// exp(-u^2) = 2^(-r^2):
//     r2 = max(|dx|^2, eps^2); rinv = rsqrt(r2); f = rinv^3 * g(u)
//     g(u) = exp(-u^2) * Q(u),  Q = 1 + u^2 + q3 u^3 + ... + q10 u^10   (tools/fit_gfactor.py)
// = 21 FP32-pipe instructions, 2 FMNMX and 2 MUFU (RSQ, EX2) per pair (see pair_packed).  The packed variant issues
// the FP32 work as FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2, new on sm_100) on PAIRS OF TARGETS, which
// halves the issue slots the FP32 pipe needs and leaves them to MUFU / FMNMX / LDS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace p2p {

constexpr int kStages = 2;
constexpr int kPolyTerms = 9;          // R(v) = c[0] + c[1] v + ... + c[8] v^8
constexpr int kFarDegree = 6;          // H(t) = cf[0] + ... + cf[6] t^6, t = 1 / (w + far_shift)
constexpr int kFarTerms = kFarDegree + 1;
constexpr unsigned kClassBit = 0x80000000u;   // set on NEAR columns (csr_pack.cuh)

struct KernelParams {
    const int4* part;        // fixed-point positions + mass bits, local then ghost particles
    const int2* leaf;        // {first particle, count}: local leaves then ghost leaves
    const long long* row_ptr;  // [nrow + 1]
    const int* col;          // source leaf ids, ascending within a row
    float4* acc;             // per local particle, accumulated into
    unsigned int* counter;   // dynamic row scheduler
    const unsigned int* n_active;   // number of rows that have work (written by the list packing)
    const int* row_order;           // those rows, fullest target leaves first (see csr_pack.cuh)
    int nrow;
    float k_fix;             // fixed-point step in kernel length units (box / 2^32 / unit)
    float eps2;              // (eps / unit)^2
    float neps2;             // -eps2
    float c[kPolyTerms];     // q[k+2] / log2(e)^((k+2)/2): coefficients in the kernel's length unit
    float out_scale;         // mass / unit^2
    float far_coord;         // coordinate offset that makes a dummy source contribute exactly 0
    // far / near split of every CSR row (csr_pack.cuh): the first row_mid[row] columns are FAR source leaves (every particle
    // pair of the leaf pair at u >= P2P_U_FAR), the rest NEAR ones; columns carry the class in bit 31
    const int* row_mid;
    int block_leaves;        // source leaves per summation block of a class
    unsigned int* err;       // sticky error flag (a source leaf larger than the stage was skipped)
    int rows_per_warp;       // 0: persistent warps; k: a warp retires once it has run for k x 2^17 cycles (second-generation kernel)
    int persist_blocks;      // ... except the warps of the last persist_blocks blocks of the grid, which run until no row is left
    float cf[kFarTerms];     // far-field polynomial H(t), t = 1 / (r'^2 + far_shift) (tools/fit_gfactor.py)
    float far_shift;         // shift in the kernel's length unit, far_s0 = nkappa * -far_shift (the shift of the fit)
    float far_s0;
    // Second-generation kernel: the length unit is extent / 2^p (p integer), so that k_fix is a POWER OF TWO and the
    // conversion of a fixed-point offset costs one rounding (int -> float) instead of two; exp(-u^2) = 2^(-kappa r'^2) with
    // kappa in [1/2, 2) absorbed into the polynomial coefficients and into the multiply that negates the exponent.
    float nkappa;            // -kappa
    float u_scale;           // u = r' u_scale (error-budget variant only)
    const int4* lbounds;     // LeafBounds of the leaves (csr_pack.cuh) or nullptr: the row's reference point is the CENTRE of its
                             // targets' bounds (offsets half as large as from the first target: half the rounding of the separations)
};

// MUFU.RSQ / MUFU.EX2.  The file is compiled with --use_fast_math, under which rsqrtf / exp2f ARE the bare
// rsqrt.approx.ftz / ex2.approx.ftz instructions; unlike inline asm they let ptxas fold an operand negation.
__device__ __forceinline__ float rsqrt_approx(float x) { return rsqrtf(x); }
__device__ __forceinline__ float ex2_approx(float x) { return exp2f(x); }
// min(-q, neps): kept as inline PTX so that the front end cannot rewrite it into -(max(q, eps)), whose
// negation would land on the FP32 pipe; ptxas folds the neg into the FMNMX source modifier (ALU pipe).
__device__ __forceinline__ float neg_min(float q, float neps) {
    float y;
    asm("{\n.reg .f32 t;\nneg.ftz.f32 t, %1;\nmin.ftz.f32 %0, t, %2;\n}" : "=f"(y) : "f"(q), "f"(neps));
    return y;
}

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- pair arithmetic ---------------------------------------------------------------------------
// t holds the NEGATED target coordinates.
template <bool TRUNC>
__device__ __forceinline__ void pair_scalar(const KernelParams& P, float sx, float sy, float sz, float tx, float ty,
                                            float tz, float& ax, float& ay, float& az) {
    const float dx = sx + tx, dy = sy + ty, dz = sz + tz;
    const float q2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float r2 = fmaxf(q2, P.eps2);
    const float rinv = rsqrt_approx(r2);
    float f;
    if (TRUNC) {
        // same formulation as pair_packed<POLY = 1>: f = (e rinv) (rinv^2 + E(w) + v O(w))
        const float e = ex2_approx(neg_min(q2, P.neps2));
        const float v = r2 * rinv;
        float E = P.c[8], O = P.c[7];
        E = fmaf(E, r2, P.c[6]); O = fmaf(O, r2, P.c[5]);
        E = fmaf(E, r2, P.c[4]); O = fmaf(O, r2, P.c[3]);
        E = fmaf(E, r2, P.c[2]); O = fmaf(O, r2, P.c[1]);
        E = fmaf(E, r2, fmaf(rinv, rinv, P.c[0]));
        f = (e * rinv) * fmaf(v, O, E);
    } else {
        f = (rinv * rinv) * rinv;
    }
    ax = fmaf(dx, f, ax);
    ay = fmaf(dy, f, ay);
    az = fmaf(dz, f, az);
}

// two targets at once with the sm_100a packed FP32 instructions (FFMA2 / FMUL2 / FADD2).
//
// With Q(u) = 1 + v^2 R(v) (v = r in kernel units, R = c0 + c1 v + ... + c8 v^8) and v * rinv = 1:
//     f = rinv^3 e Q = (e rinv) (rinv^2 + R(v)),      R(v) = E(w) + v O(w),  w = r^2
// POLY = 1 (default): E and O are two independent Horner chains in w = r2 (no extra multiply: v^2 = r2 up
//   to the rsqrt rounding), rinv^2 enters as the constant term of E through one FFMA2 (rinv*rinv + c0), and
//   the sign MUFU.EX2 needs is produced on the ALU pipe by an FMNMX with negated operands (MUFU has no
//   negate modifier on sm_100; a multiply by -1 would cost an FP32-pipe slot).
//   FP32-pipe instructions per packed pair-op: 3 (dx) + 3 (r2) + 1 (v) + 3 (E) + 3 (O) + 1 (c0 + rinv^2)
//   + 1 (E last) + 1 (T) + 2 (e rinv, f) + 3 (acc) = 21; ALU pipe: 4 FMNMX; XU pipe: 4 MUFU.
// POLY = 0: the straightforward form (23 FP32-pipe instructions, Horner in v), kept for the sweeps.
template <bool TRUNC, int POLY = 1, int NEGALU = 0>
__device__ __forceinline__ void pair_packed(const KernelParams& P, float2 sx, float2 sy, float2 sz, float2 tx, float2 ty,
                                            float2 tz, float2& ax, float2& ay, float2& az) {
    const float2 dx = __fadd2_rn(sx, tx), dy = __fadd2_rn(sy, ty), dz = __fadd2_rn(sz, tz);
    const float2 q2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    float2 r2;
    r2.x = fmaxf(q2.x, P.eps2);
    r2.y = fmaxf(q2.y, P.eps2);
    const float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    float2 f;
    if (TRUNC && POLY == 1) {
        // exp(-u^2) = 2^(-r2) in the kernel's length unit; -max(r2, eps2) = min(-r2, -eps2) is one FMNMX
        float2 e;
        if (NEGALU) {
            e = make_float2(ex2_approx(neg_min(q2.x, P.neps2)), ex2_approx(neg_min(q2.y, P.neps2)));
        } else {
            const float2 a = __fmul2_rn(r2, make_float2(-1.f, -1.f));
            e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        }
        const float2 v = __fmul2_rn(r2, rinv);
        float2 E = make_float2(P.c[8], P.c[8]), O = make_float2(P.c[7], P.c[7]);
        E = __ffma2_rn(E, r2, make_float2(P.c[6], P.c[6]));
        O = __ffma2_rn(O, r2, make_float2(P.c[5], P.c[5]));
        E = __ffma2_rn(E, r2, make_float2(P.c[4], P.c[4]));
        O = __ffma2_rn(O, r2, make_float2(P.c[3], P.c[3]));
        E = __ffma2_rn(E, r2, make_float2(P.c[2], P.c[2]));
        O = __ffma2_rn(O, r2, make_float2(P.c[1], P.c[1]));
        const float2 c0r = __ffma2_rn(rinv, rinv, make_float2(P.c[0], P.c[0]));   // c0 + rinv^2
        E = __ffma2_rn(E, r2, c0r);
        const float2 T = __ffma2_rn(v, O, E);                                     // rinv^2 + R(v)
        f = __fmul2_rn(__fmul2_rn(e, rinv), T);
    } else if (TRUNC) {
        const float2 a = __fmul2_rn(r2, make_float2(-1.f, -1.f));
        const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        const float2 v = __fmul2_rn(r2, rinv);
        float2 R = make_float2(P.c[8], P.c[8]);
#pragma unroll
        for (int k = 7; k >= 0; k--) R = __ffma2_rn(R, v, make_float2(P.c[k], P.c[k]));
        const float2 S = __ffma2_rn(v, R, rinv);
        f = __fmul2_rn(__fmul2_rn(__fmul2_rn(rinv, rinv), e), S);
    } else {
        f = __fmul2_rn(__fmul2_rn(rinv, rinv), rinv);
    }
    ax = __ffma2_rn(dx, f, ax);
    ay = __ffma2_rn(dy, f, ay);
    az = __ffma2_rn(dz, f, az);
}

// ---- per-warp shared state ---------------------------------------------------------------------
struct alignas(16) TargetPair {     // negated, relative coordinates of targets (2p, 2p+1)
    float4 xy;                      // {-x0, -x1, -y0, -y1}
    float2 z;                       // {-z0, -z1}
    float2 pad;
};

template <int TT, int STAGE>
struct alignas(128) WarpSmem {
    int4 stage[kStages][STAGE];
    TargetPair tgt[TT / 2];
    float4 out[TT];
    uint64_t full[kStages];
};

// Issues the bulk copies of the next chunk of source leaves of the current row into `stage`.
// Returns the number of particles that will land (warp-uniform) and advances e.
template <int STAGE>
__device__ __forceinline__ int issue_chunk(const KernelParams& P, int4* stage, uint64_t* bar, long long& e,
                                           long long e_end, int lane) {
    int cnt = 0, start = 0;
    if (e + lane < e_end) {
        const int s = (int)((unsigned)__ldg(P.col + e + lane) & ~kClassBit);
        const int2 ld = __ldg(P.leaf + s);
        start = ld.x;
        cnt = ld.y;
    }
    int incl = cnt;                                   // inclusive scan over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const long long remaining = e_end - e;
    const int nl = remaining < 32 ? (int)remaining : 32;
    const unsigned fit = __ballot_sync(0xffffffffu, lane < nl && incl <= STAGE);
    const int nfit = __popc(fit);                     // leaves are taken in order, so `fit` is a prefix mask
    int total = __shfl_sync(0xffffffffu, incl, nfit > 0 ? nfit - 1 : 0);
    if (nfit == 0) {
        // the next source leaf alone exceeds the stage (only unvalidated device-side ghost tables can do that): skip it
        // and raise the flag the host reads at its next synchronisation point, instead of spinning on it forever
        total = 0;
        if (lane == 0) atomicOr(P.err, 1u);
        e += 1;
    }
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)total * 16u);
    __syncwarp();
    // Source leaves that follow each other in the particle array (consecutive leaf ids of one tree do)
    // are merged into ONE bulk copy: the destination is contiguous anyway, and every copy costs ~10
    // serialised instructions (UBLKCP takes uniform operands, so the lanes issue one after the other).
    const int prev_end = __shfl_up_sync(0xffffffffu, start + cnt, 1);
    const bool in_chunk = lane < nfit;
    const bool head = in_chunk && (lane == 0 || start != prev_end);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    // last lane of my run: the lane before the next head (or the last lane of the chunk)
    const unsigned later = lane >= 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
    const int last = later ? (__ffs(later) - 2) : (nfit - 1);
    const int run_end = __shfl_sync(0xffffffffu, incl, (last < 0 ? 0 : last) & 31);
    if (head) {
        const int run = run_end - (incl - cnt);
        if (run > 0) bulk_g2s(stage + (incl - cnt), P.part + start, (uint32_t)run * 16u, bar);
    }
    e += nfit;
    return total;
}

// TT     targets per pass (accumulators in registers); rows with more targets take several passes
// NSRC   sources per lane per slice (1 or 2)
// STAGE  particles per staging buffer
template <int TT, int NSRC, int STAGE, bool TRUNC, bool PACKED, int MINB, int POLY = 1>
__global__ void __launch_bounds__(128, MINB) p2p_rows_kernel(const KernelParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem<TT, STAGE>& S = reinterpret_cast<WarpSmem<TT, STAGE>*>(smem_raw)[wid];
    constexpr int SLICE = 32 * NSRC;

    if (lane == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
    uint32_t phase0 = 0, phase1 = 0;
    // row schedule: only rows that have tasks, grouped by target occupancy (see csr_pack.cuh)
    const unsigned int n_active = __ldg(P.n_active);

    for (;;) {
        int row = -1;
        if (lane == 0) {
            const unsigned int idx = atomicAdd(P.counter, 1u);
            if (idx < n_active) row = __ldg(P.row_order + idx);
        }
        row = __shfl_sync(0xffffffffu, row, 0);
        if (row < 0) break;
        const int2 tl = __ldg(P.leaf + row);
        const int nt_all = tl.y;
        const long long e_begin = __ldg(P.row_ptr + row);
        const long long e_end = __ldg(P.row_ptr + row + 1);
        if (nt_all <= 0 || e_begin >= e_end) continue;
        // reference point of the row: its first target particle (fixed-point)
        const int4 c4 = __ldg(P.part + tl.x);

        for (int t0 = 0; t0 < nt_all; t0 += TT) {
            const int nt = min(TT, nt_all - t0);
            // targets -> shared: negated, relative to c4; padding slots repeat the first target
            {
                const int4 t4 = __ldg(P.part + tl.x + t0 + (lane < nt ? lane : 0));
                const float x = -(float)(t4.x - c4.x) * P.k_fix, y = -(float)(t4.y - c4.y) * P.k_fix,
                            z = -(float)(t4.z - c4.z) * P.k_fix;
                if (lane < TT) {
                    float* base = reinterpret_cast<float*>(&S.tgt[lane >> 1]);
                    base[0 + (lane & 1)] = x;
                    base[2 + (lane & 1)] = y;
                    base[4 + (lane & 1)] = z;
                }
            }
            __syncwarp();

            float ax[TT], ay[TT], az[TT];
#pragma unroll
            for (int j = 0; j < TT; j++) ax[j] = ay[j] = az[j] = 0.f;

            // per-lane sources of the current slice (relative FP32); dummy = far away along x
            float sx[NSRC], sy[NSRC], sz[NSRC];
#pragma unroll
            for (int q = 0; q < NSRC; q++) { sx[q] = P.far_coord; sy[q] = 0.f; sz[q] = 0.f; }
            int have = 0;                                 // sources [0, have) of the slice are already in registers

            auto load_source = [&](const int4* buf, int idx, int q) {
                const int4 s4 = buf[idx];
                sx[q] = (float)(s4.x - c4.x) * P.k_fix;
                sy[q] = (float)(s4.y - c4.y) * P.k_fix;
                sz[q] = (float)(s4.z - c4.z) * P.k_fix;
            };

            // One slice (SLICE sources, NSRC per lane) against the first 2*K targets of the pass.
            // K is a compile-time constant so that the body is straight-line code: no per-target
            // branches, constants stay in (uniform) registers and the scheduler can interleave the
            // independent target-pair chains.
            auto slice_body = [&](auto kc) {
                constexpr int K = decltype(kc)::value;
#pragma unroll
                for (int p = 0; p < K; p++) {
                    const float4 txy = S.tgt[p].xy;
                    const float2 tz = S.tgt[p].z;
                    if (PACKED) {
                        float2 x2 = make_float2(ax[2 * p], ax[2 * p + 1]), y2 = make_float2(ay[2 * p], ay[2 * p + 1]),
                               z2 = make_float2(az[2 * p], az[2 * p + 1]);
#pragma unroll
                        for (int q = 0; q < NSRC; q++)
                            pair_packed<TRUNC, (POLY & 1), (POLY >> 1)>(P, make_float2(sx[q], sx[q]), make_float2(sy[q], sy[q]), make_float2(sz[q], sz[q]),
                                               make_float2(txy.x, txy.y), make_float2(txy.z, txy.w), tz, x2, y2, z2);
                        ax[2 * p] = x2.x; ax[2 * p + 1] = x2.y;
                        ay[2 * p] = y2.x; ay[2 * p + 1] = y2.y;
                        az[2 * p] = z2.x; az[2 * p + 1] = z2.y;
                    } else {
#pragma unroll
                        for (int q = 0; q < NSRC; q++) {
                            pair_scalar<TRUNC>(P, sx[q], sy[q], sz[q], txy.x, txy.z, tz.x, ax[2 * p], ay[2 * p], az[2 * p]);
                            pair_scalar<TRUNC>(P, sx[q], sy[q], sz[q], txy.y, txy.w, tz.y, ax[2 * p + 1], ay[2 * p + 1], az[2 * p + 1]);
                        }
                    }
                }
            };
            // The whole pass (chunk loop + slice loop) is instantiated per number of target pairs K and
            // dispatched ONCE per pass, so the slice loop calls straight-line code without any dispatch.
            auto run_pass = [&](auto kc) {
                auto compute_slice = [&]() { slice_body(kc); };
                long long e = e_begin;
                int np0 = 0, np1 = 0, cur = 0;
                fence_proxy_async();
                np0 = issue_chunk<STAGE>(P, S.stage[0], &S.full[0], e, e_end, lane);
                bool more = e < e_end;

                for (;;) {                                            // chunks of the row
                    if (more) {
                        // the other stage was fully consumed (its values are in registers) before this point
                        fence_proxy_async();
                        const int n = issue_chunk<STAGE>(P, S.stage[cur ^ 1], &S.full[cur ^ 1], e, e_end, lane);
                        if (cur) np0 = n; else np1 = n;
                    }
                    const bool last_chunk = !more;
                    more = e < e_end;
                    if (cur) { mbar_wait(&S.full[1], phase1); phase1 ^= 1u; } else { mbar_wait(&S.full[0], phase0); phase0 ^= 1u; }
                    const int4* buf = S.stage[cur];
                    const int n = cur ? np1 : np0;
                    int pos = 0;
                    for (;;) {                                        // slices; ONE call site of the slice code
                        bool run, more_slices;
                        if (have == 0 && n - pos >= SLICE) {
                            // fast path: a full slice straight from the stage, no carry bookkeeping
    #pragma unroll
                            for (int q = 0; q < NSRC; q++) load_source(buf, pos + q * 32 + lane, q);
                            pos += SLICE;
                            run = true;
                            more_slices = pos < n;
                        } else {
                            // chunk boundary: top up the carried slice / keep the leftover for the next chunk
                            const int avail = n - pos;
                            const bool full = have + avail >= SLICE;
                            const int take = full ? SLICE - have : avail;
    #pragma unroll
                            for (int q = 0; q < NSRC; q++) {
                                const int k = q * 32 + lane - have;   // position of this lane's slot in the new run
                                if (k >= 0 && k < take) load_source(buf, pos + k, q);
                            }
                            pos += take;
                            have += take;
                            run = full;
                            if (!full && last_chunk && have > 0) {    // ragged tail of the row: pad with dummies
    #pragma unroll
                                for (int q = 0; q < NSRC; q++) {
                                    if (q * 32 + lane >= have) { sx[q] = P.far_coord; sy[q] = 0.f; sz[q] = 0.f; }
                                }
                                run = true;
                            }
                            if (run) have = 0;
                            more_slices = full;
                        }
                        if (run) compute_slice();
                        if (!more_slices) break;
                    }
                    __syncwarp();
                    if (last_chunk) break;
                    cur ^= 1;
                }
            };
            switch ((nt + 1) >> 1) {                              // target pairs of this pass (warp-uniform)
#define P2P_CASE(k) case k: if constexpr (k <= TT / 2) run_pass(std::integral_constant<int, k>{}); break;
                P2P_CASE(1) P2P_CASE(2) P2P_CASE(3) P2P_CASE(4) P2P_CASE(5) P2P_CASE(6) P2P_CASE(7) P2P_CASE(8)
                P2P_CASE(9) P2P_CASE(10) P2P_CASE(11) P2P_CASE(12) P2P_CASE(13) P2P_CASE(14) P2P_CASE(15) P2P_CASE(16)
#undef P2P_CASE
                default: break;
            }

            // 32-lane reduction, once per row pass
#pragma unroll
            for (int j = 0; j < TT; j++) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    ax[j] += __shfl_xor_sync(0xffffffffu, ax[j], d);
                    ay[j] += __shfl_xor_sync(0xffffffffu, ay[j], d);
                    az[j] += __shfl_xor_sync(0xffffffffu, az[j], d);
                }
                if (lane == 0) S.out[j] = make_float4(ax[j], ay[j], az[j], 0.f);
            }
            __syncwarp();
            if (lane < nt) {
                const float4 o = S.out[lane];
                float4* dst = P.acc + tl.x + t0 + lane;
                float4 a = *dst;
                a.x = fmaf(o.x, P.out_scale, a.x);
                a.y = fmaf(o.y, P.out_scale, a.y);
                a.z = fmaf(o.z, P.out_scale, a.z);
                *dst = a;
            }
            __syncwarp();
        }
    }
}


// ================================================================================================
// Second-generation row kernel: ONE pass per row (all <= 32 targets of the leaf in registers, so every source is
// staged and converted once per row instead of once per 16 targets), and two slice bodies per row:
//   NEAR  source leaves: the full kernel (eps clamp, rsqrt, exp2, degree-10 polynomial): 22 FP32-pipe instructions per pair,
//   FAR   source leaves (tight leaf bounds at least 2 r_s P2P_U_FAR apart, 2/3 of all pairs): no clamp, no rsqrt --
//         f = 2^(-w) H(t) with t = MUFU.RCP(w + 1/2) and a degree-6 polynomial (relative error 1.3e-7): 17 instructions,
//         the shift and the negation of the exponent folded into FFMAs.
// The 32 x 3 accumulators of a lane are reduced over the warp by a transposing butterfly (31 shuffles per
// component instead of 160), which leaves target j's sum in lane j.
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool TRUNC>
__device__ __forceinline__ void pair_near2(const KernelParams& P, float2 sx, float2 sy, float2 sz, float2 tx, float2 ty,
                                           float2 tz, float2& ax, float2& ay, float2& az) {
    const float2 dx = __fadd2_rn(sx, tx), dy = __fadd2_rn(sy, ty), dz = __fadd2_rn(sz, tz);
    const float2 q2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    float2 r2;
    r2.x = fmaxf(q2.x, P.eps2);
    r2.y = fmaxf(q2.y, P.eps2);
    const float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    float2 f;
    if (TRUNC) {
        const float2 a = __fmul2_rn(r2, make_float2(P.nkappa, P.nkappa));
        const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        const float2 v = __fmul2_rn(r2, rinv);
        // (a Newton step on rinv, 3 more instructions, changed the measured force error by 2 %: MUFU.RSQ is not what limits it)
        float2 E = make_float2(P.c[8], P.c[8]), O = make_float2(P.c[7], P.c[7]);
        E = __ffma2_rn(E, r2, make_float2(P.c[6], P.c[6]));
        O = __ffma2_rn(O, r2, make_float2(P.c[5], P.c[5]));
        E = __ffma2_rn(E, r2, make_float2(P.c[4], P.c[4]));
        O = __ffma2_rn(O, r2, make_float2(P.c[3], P.c[3]));
        E = __ffma2_rn(E, r2, make_float2(P.c[2], P.c[2]));
        O = __ffma2_rn(O, r2, make_float2(P.c[1], P.c[1]));
        const float2 c0r = __ffma2_rn(rinv, rinv, make_float2(P.c[0], P.c[0]));   // c0 + rinv^2
        E = __ffma2_rn(E, r2, c0r);
        const float2 T = __ffma2_rn(v, O, E);                                     // rinv^2 + R(v)
        f = __fmul2_rn(__fmul2_rn(e, rinv), T);
    } else {
        f = __fmul2_rn(__fmul2_rn(rinv, rinv), rinv);
    }
    ax = __ffma2_rn(dx, f, ax);
    ay = __ffma2_rn(dy, f, ay);
    az = __ffma2_rn(dz, f, az);
}

__device__ __forceinline__ void pair_far2(const KernelParams& P, float2 sx, float2 sy, float2 sz, float2 tx, float2 ty, float2 tz,
                                          float2& ax, float2& ay, float2& az) {
    const float2 dx = __fadd2_rn(sx, tx), dy = __fadd2_rn(sy, ty), dz = __fadd2_rn(sz, tz);
    const float2 sh = make_float2(P.far_shift, P.far_shift);
    const float2 ws = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __ffma2_rn(dx, dx, sh)));     // w + shift
    const float2 a = __ffma2_rn(ws, make_float2(P.nkappa, P.nkappa), make_float2(P.far_s0, P.far_s0));   // -kappa w
    const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
    const float2 t = make_float2(rcp_approx(ws.x), rcp_approx(ws.y));
    float2 p = make_float2(P.cf[kFarDegree], P.cf[kFarDegree]);
#pragma unroll
    for (int k = kFarDegree - 1; k >= 0; k--) p = __ffma2_rn(p, t, make_float2(P.cf[k], P.cf[k]));
    const float2 f = __fmul2_rn(e, p);
    ax = __ffma2_rn(dx, f, ax);
    ay = __ffma2_rn(dy, f, ay);
    az = __ffma2_rn(dz, f, az);
}

// error-budget variant (development aid, tests/tools/error_distribution.py): the same FP32 separations, the force factor
// in fp64 with libm's erfc / exp -- what remains of the error is coordinate rounding and FP32 accumulation
__device__ __forceinline__ void pair_exact2(const KernelParams& P, float2 sx, float2 sy, float2 sz, float2 tx, float2 ty, float2 tz,
                                            float2& ax, float2& ay, float2& az) {
    const float2 dx = __fadd2_rn(sx, tx), dy = __fadd2_rn(sy, ty), dz = __fadd2_rn(sz, tz);
    float f[2];
    const float ddx[2] = {dx.x, dx.y}, ddy[2] = {dy.x, dy.y}, ddz[2] = {dz.x, dz.y};
    for (int k = 0; k < 2; k++) {
        const double q2 = (double)ddx[k] * ddx[k] + (double)ddy[k] * ddy[k] + (double)ddz[k] * ddz[k];
        const double r = sqrt(q2), rc = fmax(r, sqrt((double)P.eps2));
        const double u = r * (double)P.u_scale;                  // kernel length unit -> u = r / 2 r_s
        const double g = erfc(u) + 1.1283791670955126 * u * exp(-u * u);
        f[k] = (float)(q2 < 1.0e5 ? g / (rc * rc * rc) : 0.0);
    }
    const float2 ff = make_float2(f[0], f[1]);
    ax = __ffma2_rn(dx, ff, ax);
    ay = __ffma2_rn(dy, ff, ay);
    az = __ffma2_rn(dz, ff, az);
}

constexpr int kRowTargets = 32;         // = P2P_MAX_LEAF
constexpr int kBlockLeaves = 256;       // source leaves per summation block of a class (8 staging chunks)

template <int STAGE>
struct alignas(128) WarpSmem2 {
    int4 stage[kStages][STAGE];
    TargetPair tgt[kRowTargets / 2];
    uint64_t full[kStages];
};

// sum over the warp of v[j] (j = 0..31, one value per target and lane); returns target `lane`'s total
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; i++) {
            const float keep = up ? v[i + o] : v[i];
            const float send = up ? v[i] : v[i + o];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// All chunks and slices of the columns [e, e_end) of one row against the first 2 K targets, with the NEAR or the FAR body.
template <int K, bool FAR, int NSRC, int STAGE, bool TRUNC, int DBG>
__device__ __forceinline__ void run_range2(const KernelParams& P, WarpSmem2<STAGE>& S, const int4 c4, long long e, const long long e_end,
                                           const int lane, uint32_t& phase0, uint32_t& phase1, float2 (&ax)[kRowTargets / 2],
                                           float2 (&ay)[kRowTargets / 2], float2 (&az)[kRowTargets / 2]) {
    constexpr int SLICE = 32 * NSRC;
    float sx[NSRC], sy[NSRC], sz[NSRC];
#pragma unroll
    for (int q = 0; q < NSRC; q++) { sx[q] = P.far_coord; sy[q] = 0.f; sz[q] = 0.f; }
    int have = 0;                                 // sources [0, have) of the slice are already in registers
    auto load_source = [&](const int4* buf, int idx, int q) {
        const int4 s4 = buf[idx];
        sx[q] = (float)(s4.x - c4.x) * P.k_fix;
        sy[q] = (float)(s4.y - c4.y) * P.k_fix;
        sz[q] = (float)(s4.z - c4.z) * P.k_fix;
    };
    auto compute_slice = [&]() {
#pragma unroll
        for (int p = 0; p < K; p++) {
            const float4 txy = S.tgt[p].xy;
            const float2 tz = S.tgt[p].z;
#pragma unroll
            for (int q = 0; q < NSRC; q++) {
                if (DBG)
                    pair_exact2(P, make_float2(sx[q], sx[q]), make_float2(sy[q], sy[q]), make_float2(sz[q], sz[q]), make_float2(txy.x, txy.y),
                                make_float2(txy.z, txy.w), tz, ax[p], ay[p], az[p]);
                else if (FAR)
                    pair_far2(P, make_float2(sx[q], sx[q]), make_float2(sy[q], sy[q]), make_float2(sz[q], sz[q]), make_float2(txy.x, txy.y),
                              make_float2(txy.z, txy.w), tz, ax[p], ay[p], az[p]);
                else
                    pair_near2<TRUNC>(P, make_float2(sx[q], sx[q]), make_float2(sy[q], sy[q]), make_float2(sz[q], sz[q]),
                                              make_float2(txy.x, txy.y), make_float2(txy.z, txy.w), tz, ax[p], ay[p], az[p]);
            }
        }
    };
    int np0 = 0, np1 = 0, cur = 0;
    fence_proxy_async();
    np0 = issue_chunk<STAGE>(P, S.stage[0], &S.full[0], e, e_end, lane);
    bool more = e < e_end;
    for (;;) {                                            // chunks of the range
        if (more) {
            // the other stage was fully consumed (its values are in registers) before this point
            fence_proxy_async();
            const int n = issue_chunk<STAGE>(P, S.stage[cur ^ 1], &S.full[cur ^ 1], e, e_end, lane);
            if (cur) np0 = n; else np1 = n;
        }
        const bool last_chunk = !more;
        more = e < e_end;
        if (cur) { mbar_wait(&S.full[1], phase1); phase1 ^= 1u; } else { mbar_wait(&S.full[0], phase0); phase0 ^= 1u; }
        const int4* buf = S.stage[cur];
        const int n = cur ? np1 : np0;
        int pos = 0;
        for (;;) {                                        // slices; ONE call site of the slice code
            bool run, more_slices;
            if (have == 0 && n - pos >= SLICE) {
#pragma unroll
                for (int q = 0; q < NSRC; q++) load_source(buf, pos + q * 32 + lane, q);
                pos += SLICE;
                run = true;
                more_slices = pos < n;
            } else {
                // chunk boundary: top up the carried slice / keep the leftover for the next chunk
                const int avail = n - pos;
                const bool full = have + avail >= SLICE;
                const int take = full ? SLICE - have : avail;
#pragma unroll
                for (int q = 0; q < NSRC; q++) {
                    const int k = q * 32 + lane - have;
                    if (k >= 0 && k < take) load_source(buf, pos + k, q);
                }
                pos += take;
                have += take;
                run = full;
                if (!full && last_chunk && have > 0) {    // ragged tail of the range: pad with dummies
#pragma unroll
                    for (int q = 0; q < NSRC; q++) {
                        if (q * 32 + lane >= have) { sx[q] = P.far_coord; sy[q] = 0.f; sz[q] = 0.f; }
                    }
                    run = true;
                }
                if (run) have = 0;
                more_slices = full;
            }
            if (run) compute_slice();
            if (!more_slices) break;
        }
        __syncwarp();
        if (last_chunk) break;
        cur ^= 1;
    }
}

// BLOCKED: classes are summed in blocks of P.block_leaves source leaves (off in the shipped configuration: it costs 2.5 %
// and buys nothing measurable); RETIRE: the time-budgeted non-persistent mode of a multi-rank step.
// Both are compile-time: carried as run-time options they cost the plain kernel 2.6 ms of 59.5 at 256^3.
template <int NSRC, int STAGE, bool TRUNC, int MINB, int DBG = 0, bool BLOCKED = false, bool RETIRE = false>
__global__ void __launch_bounds__(128, MINB) p2p_rows2_kernel(const KernelParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem2<STAGE>& S = reinterpret_cast<WarpSmem2<STAGE>*>(smem_raw)[wid];
    constexpr int TP = kRowTargets / 2;

    if (lane == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
    uint32_t phase0 = 0, phase1 = 0;
    const unsigned int n_active = __ldg(P.n_active);
    // non-persistent mode: a warp retires once it has run for rows_per_warp x 2^17 cycles (~70 us each), so that the block
    // scheduler can hand the block's slot to a kernel of a higher-priority stream (NCCL, the halo walk and packing of a
    // multi-rank step) while this kernel runs.  The budget is TIME, not rows: the warps of a block then retire within one row
    // of each other whatever the rows cost (a fixed row count left three warps of a block idle behind the one that drew a
    // dense-clump row: 0.42 instead of 0.58 of peak on the clustered 1024^3 box).  The last persist_blocks blocks never retire,
    // so the grid finishes the schedule however many rows the budgeted blocks leave.
    const bool retiring = RETIRE && P.rows_per_warp > 0 && (int)blockIdx.x < (int)gridDim.x - P.persist_blocks;
    const long long t_retire = RETIRE ? clock64() + ((long long)P.rows_per_warp << 17) : 0;

    for (bool first = true;; first = false) {
        if (RETIRE && retiring && !first && clock64() > t_retire) break;
        int row = -1;
        if (lane == 0) {
            const unsigned int idx = atomicAdd(P.counter, 1u);
            if (idx < n_active) row = __ldg(P.row_order + idx);
        }
        row = __shfl_sync(0xffffffffu, row, 0);
        if (row < 0) break;
        const int2 tl = __ldg(P.leaf + row);
        const int nt = min(tl.y, kRowTargets);
        const long long e_begin = __ldg(P.row_ptr + row);
        const long long e_end = __ldg(P.row_ptr + row + 1);
        if (nt <= 0 || e_begin >= e_end) continue;
        const long long e_mid = TRUNC ? e_begin + __ldg(P.row_mid + row) : e_begin;
        // reference point of the row (fixed point): the centre of its targets' bounds, else its first target particle
        const int4 c4 = P.lbounds ? __ldg(P.lbounds + 2 * row) : __ldg(P.part + tl.x);
        {   // targets -> shared: negated, relative to c4; padding slots repeat the first target
            const int4 t4 = __ldg(P.part + tl.x + (lane < nt ? lane : 0));
            float* base = reinterpret_cast<float*>(&S.tgt[lane >> 1]);
            base[0 + (lane & 1)] = -(float)(t4.x - c4.x) * P.k_fix;
            base[2 + (lane & 1)] = -(float)(t4.y - c4.y) * P.k_fix;
            base[4 + (lane & 1)] = -(float)(t4.z - c4.z) * P.k_fix;
        }
        __syncwarp();

        // Summation order.  A row's sources arrive in kd order, i.e. as a spatial sweep: the running sum of a lane swings to a
        // large fraction of sum |terms| before the other side of the target cancels it, and every FP32 addition rounds relative
        // to that swing (at z = 49 the net force is 1.4 % of sum |terms|).  The FAR columns come first: their many small terms
        // are summed among themselves before the few large near terms arrive (demo box: median error 2.7e-6 -> 1.9e-6 of the
        // mean force, for free).  BLOCKED: classes longer than block_leaves source leaves (dense clumps) are consumed in blocks,
        // each reduced over the warp into the per-target totals of lane j and restarted from zero.
        float rx = 0.f, ry = 0.f, rz = 0.f;
        float2 ax[TP], ay[TP], az[TP];
#pragma unroll
        for (int j = 0; j < TP; j++) ax[j] = ay[j] = az[j] = make_float2(0.f, 0.f);
        auto flush = [&]() {
            // transposing butterfly: afterwards lane j holds target j's sums
            float v[32];
#pragma unroll
            for (int j = 0; j < TP; j++) { v[2 * j] = ax[j].x; v[2 * j + 1] = ax[j].y; }
            rx += transpose_reduce32(v, lane);
#pragma unroll
            for (int j = 0; j < TP; j++) { v[2 * j] = ay[j].x; v[2 * j + 1] = ay[j].y; }
            ry += transpose_reduce32(v, lane);
#pragma unroll
            for (int j = 0; j < TP; j++) { v[2 * j] = az[j].x; v[2 * j + 1] = az[j].y; }
            rz += transpose_reduce32(v, lane);
#pragma unroll
            for (int j = 0; j < TP; j++) ax[j] = ay[j] = az[j] = make_float2(0.f, 0.f);
        };
        auto run_class = [&](int cls, long long lo, long long hi) {
            switch ((nt + 1) >> 1) {                          // target pairs of this row (warp-uniform)
#define P2P_CASE(k)                                                                                                           \
    case k:                                                                                                                   \
        if (cls) run_range2<k, false, NSRC, STAGE, TRUNC, DBG>(P, S, c4, lo, hi, lane, phase0, phase1, ax, ay, az);           \
        else run_range2<k, TRUNC, NSRC, STAGE, TRUNC, DBG>(P, S, c4, lo, hi, lane, phase0, phase1, ax, ay, az);               \
        break;
                P2P_CASE(1) P2P_CASE(2) P2P_CASE(3) P2P_CASE(4) P2P_CASE(5) P2P_CASE(6) P2P_CASE(7) P2P_CASE(8)
                P2P_CASE(9) P2P_CASE(10) P2P_CASE(11) P2P_CASE(12) P2P_CASE(13) P2P_CASE(14) P2P_CASE(15) P2P_CASE(16)
#undef P2P_CASE
                default: break;
            }
        };
        for (int cls = 0; cls < 2; cls++) {                   // 0: far columns [e_begin, e_mid), 1: near columns [e_mid, e_end)
            const long long c_begin = cls ? e_mid : e_begin, c_end = cls ? e_end : e_mid;
            if constexpr (BLOCKED) {
                for (long long lo = c_begin; lo < c_end; lo += P.block_leaves) {
                    const long long hi = lo + P.block_leaves < c_end ? lo + P.block_leaves : c_end;
                    run_class(cls, lo, hi);
                    if (hi < c_end) flush();                  // a further block of this class follows
                }
            } else {
                if (c_begin < c_end) run_class(cls, c_begin, c_end);
            }
        }
        flush();
        if (lane < nt) {
            float4* dst = P.acc + tl.x + lane;
            float4 a = *dst;
            a.x = fmaf(rx, P.out_scale, a.x);
            a.y = fmaf(ry, P.out_scale, a.y);
            a.z = fmaf(rz, P.out_scale, a.z);
            *dst = a;
        }
        __syncwarp();
    }
}

}  // namespace p2p
