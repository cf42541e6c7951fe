// p2p_kernel.cuh -- the sm_100a near-field force kernel.
//
// Replaces ComputeP2PIndexing (1_Indexing/src/photoNs_CUDA.cu:250-387, one FP64 thread per
// (target leaf, source leaf) task, sources in a per-thread stack array) and ComputeP2PDualNaive /
// ComputeP2PSelfInteractions (2_Redundant/src/photoNs_CUDA.cu:225-309, 386-458).
//
// Work decomposition (one warp = one CSR row = one target leaf at a time, persistent warps pull
// rows from an atomic counter):
//   * the row's source leaves are streamed through a private double-buffered shared-memory ring:
//     each lane owns one source leaf of the chunk and issues ONE bulk async copy
//     (cp.async.bulk, SASS UBLKCP) of that leaf's float4 {x,y,z,m} run; completion is tracked by
//     an mbarrier transaction count, so the warp never blocks on a load it issued itself;
//   * SOURCES are spread over the 32 lanes, TARGETS are walked by an unrolled loop with the target
//     coordinates broadcast from shared memory and the per-target accumulators held in registers.
//     Lane utilisation is therefore independent of the leaf occupancy (a lane=target mapping would
//     idle 14-30 % of the lanes at the reference's leaf fill of 57-76 %); the
//     32-lane reduction happens once per row, not once per tile;
//   * partial 32-source slices are carried across chunk boundaries in registers, so only the last
//     slice of a row is ragged.
// Pair arithmetic (2_Redundant/src/photoNs_CUDA.cu:432-450 with the eps branch of
// 1_Indexing/src/photoNs_CUDA.cu:346-350 as an fmax on r^2):
//     r2 = max(|dx|^2, eps^2); rinv = rsqrt(r2); f = rinv^3 * g(u),  u = r/(2 r_s)
//     g(u) = exp(-u^2) * Q(u),  Q = 1 + u^2 + q3 u^3 + ... + q10 u^10   (tools/fit_gfactor.py)
// evaluated in units where positions are pre-scaled by a power of two s ~ 1/(2 r_s), so u = kappa*r'
// with kappa in [0.71, 1.41] folded into the coefficients.  23 FP32-pipe instructions, 1 FMNMX and
// 2 MUFU (RSQ, EX2) per pair; the packed variant issues the FP32 work as FFMA2/FMUL2/FADD2 on
// target pairs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p2p {

constexpr int kStageParticles = 512;   // capacity of one staging buffer (float4 each -> 8 KB)
constexpr int kStages = 2;
constexpr int kPolyTerms = 9;          // R(v) = c[0] + c[1] v + ... + c[8] v^8

struct KernelParams {
    const float4* part;      // scaled positions + mass, local then ghost particles
    const int2* leaf;        // {first particle, count}: local leaves then ghost leaves
    const long long* row_ptr;  // [nrow + 1]
    const int* col;          // source leaf ids, ascending within a row
    float4* acc;             // per local particle, accumulated into
    unsigned int* counter;   // dynamic row scheduler
    int nrow;
    float eps2;              // (eps * s)^2
    float nlog2e_k2;         // -log2(e) * kappa^2   (exp(-u^2) = ex2(nlog2e_k2 * r'^2)); 0 for plain
    float c[kPolyTerms];     // kappa^(k+2) * q[k+2]
    float out_scale;         // s^2 (and * mass when all masses are equal)
    float far_coord;         // coordinate offset that makes a dummy source contribute exactly 0
};

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
template <bool TRUNC, bool PERMASS>
__device__ __forceinline__ void pair_scalar(const KernelParams& P, float sx, float sy, float sz, float sm, float4 t,
                                            float& ax, float& ay, float& az) {
    // t holds the NEGATED target coordinates
    float dx = sx + t.x, dy = sy + t.y, dz = sz + t.z;
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    r2 = fmaxf(r2, P.eps2);
    float rinv = rsqrt_approx(r2);
    float rinv2 = rinv * rinv;
    float f;
    if (TRUNC) {
        float e = ex2_approx(r2 * P.nlog2e_k2);
        float v = r2 * rinv;
        float R = P.c[8];
#pragma unroll
        for (int k = 7; k >= 0; k--) R = fmaf(R, v, P.c[k]);
        float S = fmaf(v, R, rinv);          // rinv * Q(u)
        f = (rinv2 * e) * S;
    } else {
        f = rinv2 * rinv;
    }
    if (PERMASS) f *= sm;
    ax = fmaf(dx, f, ax);
    ay = fmaf(dy, f, ay);
    az = fmaf(dz, f, az);
}

// two targets at once with the sm_100a packed FP32 instructions (FFMA2 / FMUL2 / FADD2)
template <bool TRUNC, bool PERMASS>
__device__ __forceinline__ void pair_packed(const KernelParams& P, float2 sx, float2 sy, float2 sz, float2 sm,
                                            float2 tx, float2 ty, float2 tz, float2& ax, float2& ay, float2& az) {
    float2 dx = __fadd2_rn(sx, tx), dy = __fadd2_rn(sy, ty), dz = __fadd2_rn(sz, tz);
    float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    r2.x = fmaxf(r2.x, P.eps2);
    r2.y = fmaxf(r2.y, P.eps2);
    float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    float2 rinv2 = __fmul2_rn(rinv, rinv);
    float2 f;
    if (TRUNC) {
        float2 a = __fmul2_rn(r2, make_float2(P.nlog2e_k2, P.nlog2e_k2));
        float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        float2 v = __fmul2_rn(r2, rinv);
        float2 R = make_float2(P.c[8], P.c[8]);
#pragma unroll
        for (int k = 7; k >= 0; k--) R = __ffma2_rn(R, v, make_float2(P.c[k], P.c[k]));
        float2 S = __ffma2_rn(v, R, rinv);
        f = __fmul2_rn(__fmul2_rn(rinv2, e), S);
    } else {
        f = __fmul2_rn(rinv2, rinv);
    }
    if (PERMASS) f = __fmul2_rn(f, sm);
    ax = __ffma2_rn(dx, f, ax);
    ay = __ffma2_rn(dy, f, ay);
    az = __ffma2_rn(dz, f, az);
}

// ---- per-warp shared state ---------------------------------------------------------------------
template <int TT>
struct alignas(128) WarpSmem {
    float4 stage[kStages][kStageParticles];
    float4 tgt[TT];            // negated target coordinates (scalar variant)
    float2 tgt2[3][TT / 2];    // packed variant: [x|y|z][pair] = {-t(2p), -t(2p+1)}
    float4 out[TT];
    uint64_t full[kStages];
};

// Issues the bulk copies of the next chunk of source leaves of the current row into `stage`.
// Returns the number of particles that will land (warp-uniform) and advances e.
__device__ __forceinline__ int issue_chunk(const KernelParams& P, float4* stage, uint64_t* bar, long long& e,
                                           long long e_end, int lane) {
    int cnt = 0, start = 0;
    if (e + lane < e_end) {
        int s = __ldg(P.col + e + lane);
        int2 ld = __ldg(P.leaf + s);
        start = ld.x;
        cnt = ld.y;
    }
    int incl = cnt;                                   // inclusive scan over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    long long remaining = e_end - e;
    int nl = remaining < 32 ? (int)remaining : 32;
    unsigned fit = __ballot_sync(0xffffffffu, lane < nl && incl <= kStageParticles);
    int nfit = __popc(fit);                           // leaves are taken in order, so `fit` is a prefix mask
    int total = __shfl_sync(0xffffffffu, incl, nfit > 0 ? nfit - 1 : 0);
    if (nfit == 0) total = 0;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)total * 16u);
    __syncwarp();
    if (lane < nfit && cnt > 0) bulk_g2s(stage + (incl - cnt), P.part + start, (uint32_t)cnt * 16u, bar);
    e += nfit;
    return total;
}

template <int TT, bool TRUNC, bool PERMASS, bool PACKED>
__global__ void __launch_bounds__(128) p2p_rows_kernel(const KernelParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem<TT>& S = reinterpret_cast<WarpSmem<TT>*>(smem_raw)[wid];

    if (lane == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
    uint32_t phase[kStages] = {0, 0};

    for (;;) {
        int row = 0;
        if (lane == 0) row = (int)atomicAdd(P.counter, 1u);
        row = __shfl_sync(0xffffffffu, row, 0);
        if (row >= P.nrow) break;
        const int2 tl = __ldg(P.leaf + row);
        const int nt = tl.y;
        long long e = __ldg(P.row_ptr + row);
        const long long e_end = __ldg(P.row_ptr + row + 1);
        if (nt <= 0 || e >= e_end) continue;

        // targets -> shared (negated); padding targets sit on the first target, results dropped
        {
            float4 t = __ldg(P.part + tl.x + (lane < nt ? lane : 0));
            if (lane < TT) {
                S.tgt[lane] = make_float4(-t.x, -t.y, -t.z, 0.f);
                reinterpret_cast<float*>(&S.tgt2[0][0])[lane] = -t.x;
                reinterpret_cast<float*>(&S.tgt2[1][0])[lane] = -t.y;
                reinterpret_cast<float*>(&S.tgt2[2][0])[lane] = -t.z;
            }
        }
        __syncwarp();

        float ax[TT], ay[TT], az[TT];
#pragma unroll
        for (int j = 0; j < TT; j++) ax[j] = ay[j] = az[j] = 0.f;

        // a dummy source far enough that exp(-u^2) (or rinv^3 in the plain kernel) flushes to 0
        const float4 t0 = S.tgt[0];
        const float fx = P.far_coord - t0.x;
        float sx = fx, sy = -t0.y, sz = -t0.z, sm = 0.f;
        int have = 0;                                    // lanes [0, have) hold carried-over sources

        int np[kStages];
        int cur = 0;
        fence_proxy_async();
        np[0] = issue_chunk(P, S.stage[0], &S.full[0], e, e_end, lane);
        bool more = e < e_end;

        auto compute_slice = [&]() {
            if (PACKED) {
                const float2 sx2 = make_float2(sx, sx), sy2 = make_float2(sy, sy), sz2 = make_float2(sz, sz),
                             sm2 = make_float2(sm, sm);
#pragma unroll
                for (int p = 0; p < TT / 2; p++) {
                    if (2 * p < nt) {
                        float2 x2 = make_float2(ax[2 * p], ax[2 * p + 1]), y2 = make_float2(ay[2 * p], ay[2 * p + 1]),
                               z2 = make_float2(az[2 * p], az[2 * p + 1]);
                        pair_packed<TRUNC, PERMASS>(P, sx2, sy2, sz2, sm2, S.tgt2[0][p], S.tgt2[1][p], S.tgt2[2][p], x2,
                                                    y2, z2);
                        ax[2 * p] = x2.x; ax[2 * p + 1] = x2.y;
                        ay[2 * p] = y2.x; ay[2 * p + 1] = y2.y;
                        az[2 * p] = z2.x; az[2 * p + 1] = z2.y;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < TT; j++) {
                    if (j < nt) pair_scalar<TRUNC, PERMASS>(P, sx, sy, sz, sm, S.tgt[j], ax[j], ay[j], az[j]);
                }
            }
        };

        for (;;) {
            const int nxt = cur ^ 1;
            if (more) {
                // the other stage was fully read (values are in registers) before this point
                fence_proxy_async();
                np[nxt] = issue_chunk(P, S.stage[nxt], &S.full[nxt], e, e_end, lane);
            }
            const bool had_more = more;
            more = e < e_end;
            mbar_wait(&S.full[cur], phase[cur]);
            phase[cur] ^= 1u;
            const float4* buf = S.stage[cur];
            const int n = np[cur];
            int pos = 0;
            while (have + (n - pos) >= 32) {
                if (lane >= have) {
                    float4 s4 = buf[pos + lane - have];
                    sx = s4.x; sy = s4.y; sz = s4.z; sm = s4.w;
                }
                pos += 32 - have;
                have = 0;
                compute_slice();
            }
            if (lane >= have && lane - have < n - pos) {
                float4 s4 = buf[pos + lane - have];
                sx = s4.x; sy = s4.y; sz = s4.z; sm = s4.w;
            }
            have += n - pos;
            __syncwarp();
            if (!had_more) break;
            cur = nxt;
        }
        if (have > 0) {                                  // ragged tail of the row
            if (lane >= have) { sx = fx; sy = -t0.y; sz = -t0.z; sm = 0.f; }
            compute_slice();
        }

        // 32-lane reduction, once per row
#pragma unroll
        for (int j = 0; j < TT; j++) {
            if (j < nt) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    ax[j] += __shfl_xor_sync(0xffffffffu, ax[j], d);
                    ay[j] += __shfl_xor_sync(0xffffffffu, ay[j], d);
                    az[j] += __shfl_xor_sync(0xffffffffu, az[j], d);
                }
                if (lane == 0) S.out[j] = make_float4(ax[j], ay[j], az[j], 0.f);
            }
        }
        __syncwarp();
        if (lane < nt) {
            float4 o = S.out[lane];
            float4* dst = P.acc + tl.x + lane;
            float4 a = *dst;
            a.x = fmaf(o.x, P.out_scale, a.x);
            a.y = fmaf(o.y, P.out_scale, a.y);
            a.z = fmaf(o.z, P.out_scale, a.z);
            *dst = a;
        }
        __syncwarp();
    }
}

}  // namespace p2p
