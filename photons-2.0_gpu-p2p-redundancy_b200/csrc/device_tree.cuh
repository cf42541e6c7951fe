// device_tree.cuh -- kd-tree build and dual-tree walk ON THE DEVICE, with the
// reference's results bit for bit:
//   build : bksort_inplace / build_kdtree / center_kdtree / build_localtree   1_Indexing/src/fmm.c:29-263
//   walk  : acceptance :266-325, walk_task_p2p :402-534, and for the periodic images prepare_sendtree2 +
//           walk_task_p2p_ext   1_Indexing/src/remotes.c:337-446,141-317
// This translation unit is compiled with -fmad=false and without fast-math: every fp64 operation below rounds
// exactly like the reference's gcc build (x86-64, no FMA contraction).
//
// Build, level-synchronous.  All nodes of one tree level are processed by the same kernels:
//   1. split value = the reference's SEQUENTIAL fp64 mean of the split coordinate in the current particle order
//      (mean_kernel: one warp per node; large nodes use an exact parallel evaluation of the sequential sum, see
//      seq_sum_warp),
//   2. "big" flags (x > mean) and one global exclusive scan of them,
//   3. the reference's Hoare-like partition expressed in closed form: with ns = number of elements <= mean, the
//      k-th big element (from the left) of [0, ns) changes place with the k-th small element (from the right) of
//      [ns, len); nothing else moves.  The one quirk of the reference loop -- its last element is never examined,
//      so a run without any big element still sends one element to the right -- is kept.
//   4. children: a side with <= maxleaf particles becomes a leaf, the others are the nodes of the next level.
// Ids in the reference's recursion order (leaves left to right, nodes in pre-order) and the kd-cell boxes are
// assigned afterwards by one bottom-up and one top-down sweep over the levels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p2p {
namespace dt {

typedef unsigned long long ull;

// ------------------------------------------------------------------------------------------------ scan
constexpr int kTile = 2048;   // elements per block of the flag scan (256 threads x 8)

__global__ void flag_tile_sums_kernel(const unsigned char* __restrict__ flag, long long n, unsigned int* __restrict__ tile) {
    __shared__ unsigned int ws[8];
    const long long base = (long long)blockIdx.x * kTile;
    unsigned int s = 0;
    for (int k = 0; k < 8; k++) {
        long long i = base + k * 256 + threadIdx.x;
        if (i < n) s += flag[i];
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        tile[blockIdx.x] = t;
    }
}

// exclusive scan of the tile sums by one block; tile[ntile] receives the total
__global__ void flag_tile_offsets_kernel(unsigned int* __restrict__ tile, int ntile) {
    __shared__ unsigned int ws[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntile; base += 1024) {
        int i = base + threadIdx.x;
        unsigned int v = i < ntile ? tile[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int w = ws[threadIdx.x], z = w;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int y = __shfl_up_sync(0xffffffffu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            ws[threadIdx.x] = z - w;
        }
        __syncthreads();
        const unsigned int excl = carry + ws[threadIdx.x >> 5] + x - v;
        if (i < ntile) tile[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile[ntile] = carry;
}

// G[i] = number of flags in [0, i); G[n] = total
__global__ void flag_scan_apply_kernel(const unsigned char* __restrict__ flag, long long n, const unsigned int* __restrict__ tile,
                                       unsigned int* __restrict__ G) {
    __shared__ unsigned int ws[8];
    const long long base = (long long)blockIdx.x * kTile + (long long)threadIdx.x * 8;
    unsigned int f[8], s = 0;
    for (int k = 0; k < 8; k++) {
        f[k] = (base + k < n) ? flag[base + k] : 0;
        s += f[k];
    }
    unsigned int x = s;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned int off = tile[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += ws[w];
    unsigned int run = off + x - s;
    for (int k = 0; k < 8; k++) {
        if (base + k < n) G[base + k] = run;
        run += f[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) G[n] = tile[gridDim.x];
}

// max over the lanes that are active here, ONE atomic per warp (per-thread atomics on one word serialise chip-wide)
__device__ __forceinline__ void warp_atomic_max(int* addr, int v) {
    const unsigned m = __activemask();
    v = __reduce_max_sync(m, v);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1 && v > 0) atomicMax(addr, v);
}

// ------------------------------------------------------------------------------------------------ build
struct BuildArrays {
    // particles in the current order (SoA) and their original indices
    double* x[3];
    int* perm;
    int* seg;            // temp node that owns position i at the current level, -1 once it is inside a leaf
    int* seg_next;
    unsigned char* flag;
    unsigned int* G;     // exclusive scan of flag, [npart + 1]
    int* slot;           // right-zone small elements by rank from the right
    // temp nodes, breadth first
    int* t_start;
    int* t_len;
    int* t_parent;       // parent temp id * 2 + side, -1 for the root
    int* t_np0;
    double* t_split;
    int* t_child;        // [2]: temp id of a child node, or -1 - (its particle count) for a leaf
    int* t_nleaf;        // leaves / nodes of the subtree
    int* t_nnode;
    int* t_id;           // final node id (pre-order) and first leaf id of the subtree
    int* t_leafbase;
    double* t_lo;        // [3] kd cell of the node
    double* t_hi;
};

// The reference's sequential sum  s = (((x0 + x1) + x2) + ...)  of one warp's run, evaluated exactly.
// Plain version: lanes load 32 values, every lane folds them in order (uniform result).
__device__ __forceinline__ double seq_sum_warp_plain(const double* __restrict__ x, long long start, int len) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int base = 0; base < len; base += 32) {
        const double v = (base + lane < len) ? x[start + base + lane] : 0.0;
        const int m = min(32, len - base);
        for (int k = 0; k < m; k++) s += __shfl_sync(0xffffffffu, v, k);
    }
    return s;
}

// Exact parallel evaluation of the same sequential sum for long runs of non-negative values.
// While the running sum S stays inside one binade [2^e, 2^(e+1)), S = k u with u = 2^(e-52) and k an integer in
// [2^52, 2^53), and  fl(S + x) = (k + q + c) u  where x = (q + r) u, q integer, 0 <= r < 1, and the round-to-
// nearest-even carry c is 1 if r > 1/2, 0 if r < 1/2 and the parity of k + q if r = 1/2.  The only state that
// crosses an addition besides the sum is therefore ONE BIT (the parity of k), so a run of additions is a
// two-state transducer {parity in -> (integer increment, parity out)} and transducers compose associatively:
// each lane folds 8 consecutive values for both input parities, a warp scan composes the 32 lanes, and the tile
// of 256 values costs O(8 + log 32) steps instead of 256 dependent additions.  A lane whose values would carry
// the sum out of the binade (or that are negative / not finite / larger than the binade) stops the tile: the
// lanes before it are applied, its own 8 values are added natively, and the next tile starts behind it with the
// new binade.  Every path produces the bits of the sequential loop.
__device__ __forceinline__ double seq_sum_warp(const double* __restrict__ x, long long start, int len) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double S = 0.0;
    int pos = 0;
    while (pos < len) {
        const int e = (int)((__double_as_longlong(S) >> 52) & 0x7ff) - 1023;       // binade of S (S >= 0)
        const bool okS = S > 0.0 && e >= -900 && e <= 900;
        const int my = pos + lane * 8;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (my + k < len) ? x[start + my + k] : 0.0;
        // fold my 8 values for both input parities
        long long d0 = 0, d1 = 0;
        int p0 = 0, p1 = 1;
        bool hard = !okS;
        const double scale = okS ? __longlong_as_double((long long)(1023 + 52 - e) << 52) : 1.0;   // 2^(52-e)
        const double top = okS ? __longlong_as_double((long long)(1023 + e + 1) << 52) : 0.0;      // 2^(e+1)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double xv = v[k];
            if (!(xv >= 0.0 && xv < top)) hard = true;
            const double y = xv * scale;                 // exact: power-of-two scaling, no underflow for xv >= 0 normal or 0
            const double qf = floor(y);
            const double r = y - qf;                     // exact
            const long long q = (long long)qf;
            const int gt = r > 0.5, tie = r == 0.5;
            int t0 = p0 ^ (int)(q & 1), t1 = p1 ^ (int)(q & 1);
            const int c0 = gt | (tie & t0), c1 = gt | (tie & t1);
            d0 += q + c0;
            d1 += q + c1;
            p0 = t0 ^ c0;
            p1 = t1 ^ c1;
        }
        // a denormal input would lose bits in the scaling only if the product were denormal too; xv * 2^(52-e)
        // with e <= 900 is >= xv, so it is exact.
        // inclusive warp scan of the transducers (compose: first A then B)
        long long a0 = d0, a1 = d1;
        int q0 = p0, q1 = p1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long b0 = __shfl_up_sync(full, a0, o), b1 = __shfl_up_sync(full, a1, o);
            const int r0 = __shfl_up_sync(full, q0, o), r1 = __shfl_up_sync(full, q1, o);
            if (lane >= o) {
                // earlier lanes (b, r) first, then mine (a, q)
                const long long n0 = b0 + (r0 ? a1 : a0), n1 = b1 + (r1 ? a1 : a0);
                const int m0 = r0 ? q1 : q0, m1 = r1 ? q1 : q0;
                a0 = n0; a1 = n1; q0 = m0; q1 = m1;
            }
        }
        const long long kS = okS ? ((__double_as_longlong(S) & 0xfffffffffffffLL) | (1LL << 52)) : 0;   // integer mantissa
        const int parS = (int)(kS & 1);
        const long long incl = parS ? a1 : a0;                       // increment after my lane
        const bool cross = hard || (kS + incl >= (1LL << 53));
        const unsigned bad = __ballot_sync(full, cross);
        if (bad == 0) {
            const long long tot = __shfl_sync(full, incl, 31);
            S = __longlong_as_double((long long)(1023 + e - 52) << 52) * (double)(kS + tot);          // (k + tot) u, exact
            pos += 256;
            continue;
        }
        const int f = __ffs(bad) - 1;
        if (f > 0) {
            const long long before = __shfl_sync(full, incl, f - 1);
            S = __longlong_as_double((long long)(1023 + e - 52) << 52) * (double)(kS + before);
        }
        // lane f's own values, natively and in order
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double xv = __shfl_sync(full, v[k], f);
            S += xv;                                                   // beyond the run: + 0.0
        }
        pos += 8 * (f + 1);
    }
    return S;
}

constexpr int kSeqPlainMax = 256;    // runs up to this length use the plain fold
constexpr int kBlockWarps = 32;      // warps of the block-per-node variant (tile = 32 x 256 values)

// one warp per node of the level: split value.  Nodes longer than skip_above are left to mean_block_kernel.
__global__ void mean_kernel(BuildArrays A, int lvl_begin, int lvl_count, int dir, int plain_max, int skip_above) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const double* __restrict__ X = A.x[dir];
    for (int n = w; n < lvl_count; n += nw) {
        const int t = lvl_begin + n;
        const long long start = A.t_start[t];
        const int len = A.t_len[t];
        if (len > skip_above) continue;
        double split = 0.0;
        if (len == 2) {
            split = 0.5 * (X[start] + X[start + 1]);
        } else if (len > 2) {
            const double s = len <= plain_max ? seq_sum_warp_plain(X, start, len) : seq_sum_warp(X, start, len);
            split = s / (double)len;
        }
        if ((threadIdx.x & 31) == 0) A.t_split[t] = split;
    }
}

// Shared memory of the block-cooperative transducer evaluation.
struct BlockSum {
    long long a0[kBlockWarps], a1[kBlockWarps], pre[kBlockWarps];
    int q0[kBlockWarps], q1[kBlockWarps], par[kBlockWarps], bad[kBlockWarps];
    long long tot;
    double v[8];
};

// fold 8 values for both input parities (binade exponent e of the running sum); hard: a value outside [0, 2^(e+1))
__device__ __forceinline__ void fold8(const double v[8], bool okS, int e, long long& d0, long long& d1, int& p0, int& p1, bool& hard) {
    d0 = 0; d1 = 0; p0 = 0; p1 = 1;
    hard = !okS;
    const double scale = okS ? __longlong_as_double((long long)(1023 + 52 - e) << 52) : 1.0;
    const double top = okS ? __longlong_as_double((long long)(1023 + e + 1) << 52) : 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double xv = v[k];
        if (!(xv >= 0.0 && xv < top)) hard = true;
        const double y = xv * scale;
        const double qf = floor(y);
        const double r = y - qf;
        const long long q = (long long)qf;
        const int gt = r > 0.5, tie = r == 0.5;
        int t0 = p0 ^ (int)(q & 1), t1 = p1 ^ (int)(q & 1);
        const int c0 = gt | (tie & t0), c1 = gt | (tie & t1);
        d0 += q + c0;
        d1 += q + c1;
        p0 = t0 ^ c0;
        p1 = t1 ^ c1;
    }
}

// inclusive warp scan of transducers (earlier lanes first)
__device__ __forceinline__ void scan_transducers(long long& a0, long long& a1, int& q0, int& q1) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long b0 = __shfl_up_sync(full, a0, o), b1 = __shfl_up_sync(full, a1, o);
        const int r0 = __shfl_up_sync(full, q0, o), r1 = __shfl_up_sync(full, q1, o);
        if (lane >= o) {
            const long long n0 = b0 + (r0 ? a1 : a0), n1 = b1 + (r1 ? a1 : a0);
            const int m0 = r0 ? q1 : q0, m1 = r1 ? q1 : q0;
            a0 = n0; a1 = n1; q0 = m0; q1 = m1;
        }
    }
}

// The transducer evaluation with a whole block (32 warps, tile = 8192 values): continues the sequential sum S over
// X[start + lo .. start + hi); every thread folds 8 values, warps scan, warp 0 scans the warp summaries, and the
// first thread whose values leave the binade ends the tile exactly as in seq_sum_warp.  Uniform result.
__device__ double block_seq_sum(const double* __restrict__ X, long long start, int lo, int hi, double S, BlockSum& sm) {
    constexpr int NW = kBlockWarps;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int pos = lo;
    while (pos < hi) {
        const int e = (int)((__double_as_longlong(S) >> 52) & 0x7ff) - 1023;
        const bool okS = S > 0.0 && e >= -900 && e <= 900;
        const int my = pos + tid * 8;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (my + k < hi) ? X[start + my + k] : 0.0;
        long long a0, a1;
        int q0, q1;
        bool hard;
        fold8(v, okS, e, a0, a1, q0, q1, hard);
        scan_transducers(a0, a1, q0, q1);
        if (lane == 31) { sm.a0[w] = a0; sm.a1[w] = a1; sm.q0[w] = q0; sm.q1[w] = q1; }
        const long long kS = okS ? ((__double_as_longlong(S) & 0xfffffffffffffLL) | (1LL << 52)) : 0;
        const int parS = (int)(kS & 1);
        __syncthreads();
        if (w == 0) {
            long long A0 = sm.a0[lane], A1 = sm.a1[lane];
            int Q0 = sm.q0[lane], Q1 = sm.q1[lane];
            scan_transducers(A0, A1, Q0, Q1);
            const long long inc = parS ? A1 : A0;
            const int par = parS ? Q1 : Q0;
            long long ex = __shfl_up_sync(full, inc, 1);
            int exp_ = __shfl_up_sync(full, par, 1);
            if (lane == 0) { ex = 0; exp_ = parS; }
            sm.pre[lane] = ex;
            sm.par[lane] = exp_;
        }
        __syncthreads();
        const long long incl = sm.pre[w] + (sm.par[w] ? a1 : a0);
        const bool cross = hard || (kS + incl >= (1LL << 53));
        const unsigned bad = __ballot_sync(full, cross);
        if (lane == 0) sm.bad[w] = bad ? (w * 32 + __ffs(bad) - 1) : 0x7fffffff;
        __syncthreads();
        int f = sm.bad[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) f = min(f, __shfl_xor_sync(full, f, o));
        const double u = __longlong_as_double((long long)(1023 + (okS ? e : 0) - 52) << 52);
        if (f == 0x7fffffff) {
            if (tid == NW * 32 - 1) sm.tot = incl;
            __syncthreads();
            S = u * (double)(kS + sm.tot);
            pos += NW * 256;
        } else {
            if (tid == f - 1) sm.tot = incl;
            if (tid == f) {
#pragma unroll
                for (int k = 0; k < 8; k++) sm.v[k] = v[k];
            }
            __syncthreads();
            if (f > 0) S = u * (double)(kS + sm.tot);
#pragma unroll
            for (int k = 0; k < 8; k++) S += sm.v[k];
            pos += 8 * (f + 1);
        }
        __syncthreads();
    }
    return S;
}

// one block per node (nodes with min_len < len <= max_len)
__global__ void __launch_bounds__(kBlockWarps * 32) mean_block_kernel(BuildArrays A, int lvl_begin, int lvl_count, int dir, int min_len,
                                                                      int max_len) {
    __shared__ BlockSum sm;
    const double* __restrict__ X = A.x[dir];
    for (int n = blockIdx.x; n < lvl_count; n += gridDim.x) {
        const int t = lvl_begin + n;
        const int len = A.t_len[t];
        if (len <= min_len || len > max_len) continue;
        const double S = block_seq_sum(X, A.t_start[t], 0, len, 0.0, sm);
        if (threadIdx.x == 0) A.t_split[t] = S / (double)len;
    }
}

// ---- speculative evaluation for the longest nodes: all SMs work on ONE sequential sum ---------------------------
// The node is cut into chunks of 8192 values.  A plain (tree-order) sum per chunk and a prefix over the chunks give
// the running sum at every chunk start to ~1e-13 -- enough to know its BINADE unless it is within 1e-6 of a power
// of two.  Every such "safe" chunk then computes its transducer {parity in -> (increment, parity out)} for that
// binade independently, on any SM; one block per node finally chains the chunk transducers in order, checking for
// every chunk that the exact running sum really is in the predicted binade and stays in it, and evaluates the few
// chunks where that fails (binade crossings) cooperatively from the exact running sum.  Bits identical to the loop.
constexpr int kChunk = kBlockWarps * 256;

struct SpecArrays {
    int* choff;          // [nodes of the level + 1] first chunk of each node (0 chunks for nodes not treated here)
    double* approx;      // [chunk] plain sum, then (prefix kernel) approximate running sum at the chunk start
    double* approx_end;
    long long* inc;      // [chunk][2]
    int* meta;           // [chunk] bit 0 safe, bits 1-2 parity out for parity in 0 / 1, bits 8.. binade exponent + 2048
};

__global__ void chunk_offsets_kernel(BuildArrays A, SpecArrays Sp, int lvl_begin, int lvl_count, int min_len) {
    // single block; levels treated here have at most a few hundred nodes
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < lvl_count; base += blockDim.x) {
        const int n = base + threadIdx.x;
        int c = 0;
        if (n < lvl_count) { const int len = A.t_len[lvl_begin + n]; c = len > min_len ? (len + kChunk - 1) / kChunk : 0; }
        // serial prefix by thread 0 over this batch (tiny)
        __shared__ int tmp[1024];
        tmp[threadIdx.x] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = carry;
            for (int k = 0; k < (int)blockDim.x && base + k < lvl_count; k++) { const int v = tmp[k]; tmp[k] = run; run += v; }
            carry = run;
        }
        __syncthreads();
        if (n < lvl_count) Sp.choff[n] = tmp[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) Sp.choff[lvl_count] = carry;
}

__device__ __forceinline__ int node_of_chunk(const int* __restrict__ choff, int lvl_count, int b) {
    int lo = 0, hi = lvl_count;              // last n with choff[n] <= b
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (choff[mid] <= b) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) chunk_sum_kernel(BuildArrays A, SpecArrays Sp, int lvl_begin, int lvl_count, int dir) {
    __shared__ double ws[8];
    const double* __restrict__ X = A.x[dir];
    const int total = Sp.choff[lvl_count];
    for (int b = blockIdx.x; b < total; b += gridDim.x) {
        const int n = node_of_chunk(Sp.choff, lvl_count, b);
        const int t = lvl_begin + n, c = b - Sp.choff[n];
        const long long start = A.t_start[t];
        const int len = A.t_len[t], lo = c * kChunk, hi = min(len, lo + kChunk);
        double s = 0.0;
        for (int i = lo + threadIdx.x; i < hi; i += 256) s += X[start + i];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) { double r = 0.0; for (int k = 0; k < 8; k++) r += ws[k]; Sp.approx[b] = r; }
        __syncthreads();
    }
}

// one warp per node: approximate running sum at the start / end of each of its chunks
__global__ void chunk_prefix_kernel(SpecArrays Sp, int lvl_count) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= lvl_count) return;
    const int b0 = Sp.choff[w], b1 = Sp.choff[w + 1];
    double run = 0.0;
    for (int base = b0; base < b1; base += 32) {
        const int b = base + lane;
        const double v = b < b1 ? Sp.approx[b] : 0.0;
        double x = v;
        for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (b < b1) { Sp.approx[b] = run + (x - v); Sp.approx_end[b] = run + x; }
        run += __shfl_sync(0xffffffffu, x, 31);
    }
}

__global__ void __launch_bounds__(kBlockWarps * 32) chunk_transducer_kernel(BuildArrays A, SpecArrays Sp, int lvl_begin, int lvl_count, int dir) {
    __shared__ long long s_a0[kBlockWarps], s_a1[kBlockWarps];
    __shared__ int s_q0[kBlockWarps], s_q1[kBlockWarps], s_hard;
    const double* __restrict__ X = A.x[dir];
    const int total = Sp.choff[lvl_count];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int b = blockIdx.x; b < total; b += gridDim.x) {
        const double Sa = Sp.approx[b], Sb = Sp.approx_end[b];
        const int e = (int)((__double_as_longlong(Sa) >> 52) & 0x7ff) - 1023;
        bool safe = Sa > 0.0 && e >= -900 && e <= 900;
        if (safe) {
            const double lo2 = __longlong_as_double((long long)(1023 + e) << 52), hi2 = lo2 * 2.0;
            safe = Sa > lo2 * (1.0 + 1e-6) && Sb < hi2 * (1.0 - 1e-6);
        }
        if (!safe) { if (tid == 0) Sp.meta[b] = 0; continue; }          // uniform branch: Sa, Sb are the same for the block
        const int n = node_of_chunk(Sp.choff, lvl_count, b);
        const int t = lvl_begin + n, c = b - Sp.choff[n];
        const long long start = A.t_start[t];
        const int len = A.t_len[t], lo = c * kChunk, hi = min(len, lo + kChunk);
        if (tid == 0) s_hard = 0;
        __syncthreads();
        const int my = lo + tid * 8;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (my + k < hi) ? X[start + my + k] : 0.0;
        long long a0, a1;
        int q0, q1;
        bool hard;
        fold8(v, true, e, a0, a1, q0, q1, hard);
        if (hard) s_hard = 1;
        scan_transducers(a0, a1, q0, q1);
        if (lane == 31) { s_a0[w] = a0; s_a1[w] = a1; s_q0[w] = q0; s_q1[w] = q1; }
        __syncthreads();
        if (w == 0) {
            long long A0 = s_a0[lane], A1 = s_a1[lane];
            int Q0 = s_q0[lane], Q1 = s_q1[lane];
            scan_transducers(A0, A1, Q0, Q1);
            if (lane == 31) {
                Sp.inc[2 * (size_t)b] = A0;
                Sp.inc[2 * (size_t)b + 1] = A1;
                Sp.meta[b] = s_hard ? 0 : (1 | (Q0 << 1) | (Q1 << 2) | ((e + 2048) << 8));
            }
        }
        __syncthreads();
    }
}

// one block per node: chain the chunk transducers in order from the exact running sum; unsafe or mispredicted chunks
// are evaluated cooperatively
__global__ void __launch_bounds__(kBlockWarps * 32) chunk_combine_kernel(BuildArrays A, SpecArrays Sp, int lvl_begin, int lvl_count, int dir) {
    __shared__ BlockSum sm;
    // The chain is sequential by nature (chunk b starts from the exact sum behind chunk b - 1), so what it costs is the latency
    // of one step.  The chunk records of a node are therefore staged in shared memory a block-load at a time -- read straight
    // from global memory every step waited for two dependent loads: 0.3 ms for the 2048 chunks of the 256^3 root.
    __shared__ int s_meta[kBlockWarps * 32];
    __shared__ long long s_inc[2 * kBlockWarps * 32];
    const double* __restrict__ X = A.x[dir];
    for (int n = blockIdx.x; n < lvl_count; n += gridDim.x) {
        const int b0 = Sp.choff[n], b1 = Sp.choff[n + 1];
        if (b1 == b0) continue;
        const int t = lvl_begin + n;
        const long long start = A.t_start[t];
        const int len = A.t_len[t];
        double S = 0.0;                                   // every thread carries the same value
        for (int base = b0; base < b1; base += kBlockWarps * 32) {
            const int nb = min(b1 - base, kBlockWarps * 32);
            __syncthreads();                              // the previous batch has been consumed
            if ((int)threadIdx.x < nb) {
                const int b = base + threadIdx.x;
                s_meta[threadIdx.x] = Sp.meta[b];
                s_inc[2 * threadIdx.x] = Sp.inc[2 * (size_t)b];
                s_inc[2 * threadIdx.x + 1] = Sp.inc[2 * (size_t)b + 1];
            }
            __syncthreads();
            for (int j = 0; j < nb; j++) {
                const int meta = s_meta[j];
                bool done = false;
                if (meta & 1) {
                    const int e = (meta >> 8) - 2048;
                    const int eS = (int)((__double_as_longlong(S) >> 52) & 0x7ff) - 1023;
                    if (S > 0.0 && eS == e) {
                        const long long kS = (__double_as_longlong(S) & 0xfffffffffffffLL) | (1LL << 52);
                        const long long inc = s_inc[2 * j + (int)(kS & 1)];
                        if (kS + inc < (1LL << 53)) {
                            S = __longlong_as_double((long long)(1023 + e - 52) << 52) * (double)(kS + inc);
                            done = true;
                        }
                    }
                }
                if (!done) {                              // uniform: S and meta are the same for every thread
                    const int lo = (base - b0 + j) * kChunk;
                    S = block_seq_sum(X, start, lo, min(len, lo + kChunk), S, sm);
                }
            }
        }
        if (threadIdx.x == 0) A.t_split[t] = S / (double)len;
    }
}

__global__ void flag_kernel(BuildArrays A, long long npart, int dir) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npart) return;
    const int t = A.seg[i];
    unsigned char f = 0;
    if (t >= 0) {
        const int len = A.t_len[t];
        const double* __restrict__ X = A.x[dir];
        if (len == 2) {
            const long long s = A.t_start[t];
            const bool swap = X[s] > X[s + 1];
            f = (i == s) ? swap : !swap;
        } else if (len > 2) {
            f = X[i] > A.t_split[t];
        }
    }
    A.flag[i] = f;
}

// per node: left count (with the reference's quirks) and the children
__global__ void split_kernel(BuildArrays A, int lvl_begin, int lvl_count, int maxleaf, int* __restrict__ nchild_nodes) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    const long long start = A.t_start[t];
    const int len = A.t_len[t];
    int np0;
    if (len < 2) np0 = 0;
    else if (len == 2) np0 = 1;
    else {
        const int nbig = (int)(A.G[start + len] - A.G[start]);
        np0 = nbig == 0 ? len - 1 : len - nbig;
    }
    A.t_np0[t] = np0;
    const int np1 = len - np0;
    nchild_nodes[n] = (np0 > maxleaf) + (np1 > maxleaf);
}

// exclusive scan of the per-node child counts of one level by a single block (levels hold at most a few 10^5 nodes)
__global__ void child_scan_kernel(int* __restrict__ cnt, int n, int* __restrict__ total) {
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < n ? cnt[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int wv = ws[threadIdx.x], z = wv;
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            ws[threadIdx.x] = z - wv;
        }
        __syncthreads();
        const int excl = carry + ws[threadIdx.x >> 5] + x - v;
        if (i < n) cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void children_kernel(BuildArrays A, int lvl_begin, int lvl_count, int maxleaf, const int* __restrict__ child_off,
                                int next_begin, int node_cap, int* __restrict__ longest) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    const int start = A.t_start[t], len = A.t_len[t], np0 = A.t_np0[t], np1 = len - np0;
    int next = next_begin + child_off[n];
    int c0, c1;
    if (np0 > maxleaf) {
        c0 = next++;
        if (c0 < node_cap) { A.t_start[c0] = start; A.t_len[c0] = np0; A.t_parent[c0] = 2 * t; }
    } else c0 = -1 - np0;
    if (np1 > maxleaf) {
        c1 = next++;
        if (c1 < node_cap) { A.t_start[c1] = start + np0; A.t_len[c1] = np1; A.t_parent[c1] = 2 * t + 1; }
    } else c1 = -1 - np1;
    A.t_child[2 * t] = c0;
    A.t_child[2 * t + 1] = c1;
    warp_atomic_max(longest, max(np0 > maxleaf ? np0 : 0, np1 > maxleaf ? np1 : 0));
}

// positions of the right-zone small elements by rank from the right, and the owner of every position at the next level
__global__ void slot_kernel(BuildArrays A, long long npart) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npart) return;
    const int t = A.seg[i];
    int next = -1;
    if (t >= 0) {
        const long long start = A.t_start[t];
        const int len = A.t_len[t], np0 = A.t_np0[t];
        const int rel = (int)(i - start);
        const int c = A.t_child[2 * t + (rel >= np0)];
        next = c >= 0 ? c : -1;
        if (rel >= np0 && !A.flag[i]) {
            const int cb = (int)(A.G[i] - A.G[start]);                         // big elements left of i
            const int nbig = (int)(A.G[start + len] - A.G[start]);
            const int r = (len - nbig) - (rel - cb) - 1;                       // small elements right of i
            A.slot[start + r] = (int)i;
        }
    }
    A.seg_next[i] = next;
}

__global__ void swap_kernel(BuildArrays A, long long npart) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npart) return;
    const int t = A.seg[i];
    if (t < 0 || !A.flag[i]) return;
    const long long start = A.t_start[t];
    const int np0 = A.t_np0[t];
    if (i - start >= np0) return;
    const int k = (int)(A.G[i] - A.G[start]);
    const long long j = A.slot[start + k];
    for (int d = 0; d < 3; d++) {
        const double a = A.x[d][i], b = A.x[d][j];
        A.x[d][i] = b;
        A.x[d][j] = a;
    }
    const int pa = A.perm[i], pb = A.perm[j];
    A.perm[i] = pb;
    A.perm[j] = pa;
}

// ---- node-centric level (deep levels: every node of the level holds at most kNodeMax particles) ------------------------
// One warp does for its node what the particle-wide kernels above do for all nodes together -- split mean, flags, left
// count, the pairing of the k-th big element of the left zone with the small element of the right zone that has k small
// elements to its right, the swaps -- with the flags as ballot masks in shared memory instead of a global flag array, a
// global scan and a slot array: one kernel and one read of the split coordinate per level instead of eight kernels and six
// passes over all particles.  Same closed form, same special cases (runs of 2 and < 2), same split arithmetic.
constexpr int kNodeMax = 2048;
constexpr int kNodeWarps = 4;

struct NodeSmem {
    unsigned int mask[kNodeMax / 32];        // flag bits (x > split) of the node's particles, 32 per word
    unsigned short big[kNodeMax / 2];        // left-zone big elements in index order (relative positions)
    unsigned short small_r[kNodeMax / 2];    // right-zone small elements by the number of small elements to their right
};

__global__ void __launch_bounds__(kNodeWarps * 32) node_level_kernel(BuildArrays A, int lvl_begin, int lvl_count, int dir, int maxleaf,
                                                                   int plain_max, int* __restrict__ nchild_nodes) {
    __shared__ NodeSmem sm_all[kNodeWarps];
    NodeSmem& sm = sm_all[threadIdx.x >> 5];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const double* __restrict__ X = A.x[dir];
    for (int n = w; n < lvl_count; n += nw) {
        const int t = lvl_begin + n;
        const long long start = A.t_start[t];
        const int len = A.t_len[t];
        int np0 = 0;
        double split = 0.0;
        if (len == 2) {
            const double x0 = X[start], x1 = X[start + 1];
            split = 0.5 * (x0 + x1);
            np0 = 1;
            __syncwarp();                                        // every lane has read the pair
            if (x0 > x1 && lane < 4) {                           // the two are put in ascending order
                if (lane < 3) { const double a = A.x[lane][start], b = A.x[lane][start + 1]; A.x[lane][start] = b; A.x[lane][start + 1] = a; }
                else { const int a = A.perm[start], b = A.perm[start + 1]; A.perm[start] = b; A.perm[start + 1] = a; }
            }
        } else if (len > 2) {
            const double ssum = len <= plain_max ? seq_sum_warp_plain(X, start, len) : seq_sum_warp(X, start, len);
            split = ssum / (double)len;
            // flags as ballot masks; number of big elements
            const int nword = (len + 31) >> 5;
            int nbig = 0;
            for (int c = 0; c < nword; c++) {
                const int i = c * 32 + lane;
                const bool f = i < len && X[start + i] > split;
                const unsigned m = __ballot_sync(full, f);
                if (lane == 0) sm.mask[c] = m;
                nbig += __popc(m);
            }
            __syncwarp();
            np0 = nbig == 0 ? len - 1 : len - nbig;
            if (nbig > 0) {
                // every lane walks the words with running counts (uniform), and files its own element of the word
                int big_before = 0;                              // big elements in the words before c
                const int nsmall = len - nbig;
                for (int c = 0; c < nword; c++) {
                    const unsigned m = sm.mask[c];
                    const int i = c * 32 + lane;
                    if (i < len) {
                        const int bl = big_before + __popc(m & ((1u << lane) - 1u));      // big elements left of i
                        const bool f = (m >> lane) & 1u;
                        if (i < np0) { if (f) sm.big[bl] = (unsigned short)i; }
                        else if (!f) sm.small_r[nsmall - (i - bl) - 1] = (unsigned short)i;  // i - bl small elements left of i
                    }
                    big_before += __popc(m);
                }
                __syncwarp();
                // big elements of the left zone = small elements of the right zone = np0 - (small elements of the left zone)
                int big_left = 0;
                for (int c = 0; c * 32 < np0; c++) {
                    const unsigned m = sm.mask[c];
                    const int rem = np0 - c * 32;
                    big_left += __popc(rem >= 32 ? m : (m & ((1u << rem) - 1u)));
                }
                for (int k = lane; k < big_left; k += 32) {
                    const long long i = start + sm.big[k], j = start + sm.small_r[k];
#pragma unroll
                    for (int d = 0; d < 3; d++) { const double a = A.x[d][i], b = A.x[d][j]; A.x[d][i] = b; A.x[d][j] = a; }
                    const int pa = A.perm[i], pb = A.perm[j];
                    A.perm[i] = pb; A.perm[j] = pa;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            A.t_split[t] = split;
            A.t_np0[t] = np0;
            nchild_nodes[n] = (np0 > maxleaf) + (len - np0 > maxleaf);
        }
    }
}

// ---- block-centric level (middle levels: enough nodes to fill the chip, none above kBlockNodeMax particles) ------------
// The same closed form with one BLOCK per node: every warp owns a contiguous range of the node's 32-particle words, counts
// its big elements in a first sweep (flags kept as ballot masks in shared memory), a scan over the warp counts gives every
// warp its starting rank, and a second sweep over the masks files the left-zone big elements and the right-zone small
// elements into two lists (the node's own range of the slot / seg_next arrays, which the particle-wide kernels no longer
// need once a level runs here: the later levels all run here or in node_level_kernel).  Replaces flag, the three scan
// kernels, split, slot and swap of a level; the split means come from the mean kernels as before.
constexpr int kBlockNodeMax = 131072;
constexpr int kBlockLevelThreads = 512;

__global__ void __launch_bounds__(kBlockLevelThreads) block_level_kernel(BuildArrays A, int lvl_begin, int lvl_count, int dir, int maxleaf,
                                                                       int* __restrict__ nchild_nodes) {
    constexpr int NW = kBlockLevelThreads / 32;
    __shared__ unsigned int mask[kBlockNodeMax / 32];
    __shared__ int wcnt[NW];
    __shared__ int s_big_left;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* __restrict__ X = A.x[dir];
    for (int n = blockIdx.x; n < lvl_count; n += gridDim.x) {
        const int t = lvl_begin + n;
        const long long start = A.t_start[t];
        const int len = A.t_len[t];
        int np0 = 0;
        if (len == 2) {
            const double x0 = X[start], x1 = X[start + 1];
            np0 = 1;
            __syncthreads();                                     // every thread has read the pair
            if (x0 > x1 && tid < 4) {                            // the two are put in ascending order
                if (tid < 3) { const double a = A.x[tid][start], b = A.x[tid][start + 1]; A.x[tid][start] = b; A.x[tid][start + 1] = a; }
                else { const int a = A.perm[start], b = A.perm[start + 1]; A.perm[start] = b; A.perm[start + 1] = a; }
            }
        } else if (len > 2) {
            const double split = A.t_split[t];
            const int nword = (len + 31) >> 5;
            const int w0 = (int)((long long)nword * wid / NW), w1 = (int)((long long)nword * (wid + 1) / NW);
            int cnt = 0;
            for (int c = w0; c < w1; c++) {
                const int i = c * 32 + lane;
                const bool f = i < len && X[start + i] > split;
                const unsigned m = __ballot_sync(full, f);
                if (lane == 0) mask[c] = m;
                cnt += __popc(m);
            }
            if (lane == 0) wcnt[wid] = cnt;
            if (tid == 0) s_big_left = 0;
            __syncthreads();
            int run = 0, nbig = 0;
            for (int v = 0; v < NW; v++) { const int x = wcnt[v]; if (v < wid) run += x; nbig += x; }
            np0 = nbig == 0 ? len - 1 : len - nbig;
            if (nbig > 0) {                                      // uniform over the block
                const int nsmall = len - nbig;
                int* __restrict__ big_list = A.slot + start;     // left-zone big elements in index order
                int* __restrict__ small_list = A.seg_next + start;   // right-zone small elements by the number of small elements to their right
                int my_left = 0;
                for (int c = w0; c < w1; c++) {
                    const unsigned m = mask[c];
                    const int i = c * 32 + lane;
                    if (i < len) {
                        const int bl = run + __popc(m & ((1u << lane) - 1u));      // big elements left of i
                        const bool f = (m >> lane) & 1u;
                        if (i < np0) { if (f) big_list[bl] = i; }
                        else if (!f) small_list[nsmall - (i - bl) - 1] = i;        // i - bl small elements left of i
                    }
                    const int rem = np0 - c * 32;
                    my_left += __popc(rem >= 32 ? m : (rem > 0 ? (m & ((1u << rem) - 1u)) : 0u));
                    run += __popc(m);
                }
                if (lane == 0 && my_left) atomicAdd(&s_big_left, my_left);
                __syncthreads();                                 // lists complete (global writes of this block are visible to it)
                const int big_left = s_big_left;
                for (int k = tid; k < big_left; k += kBlockLevelThreads) {
                    const long long i = start + big_list[k], j = start + small_list[k];
#pragma unroll
                    for (int d = 0; d < 3; d++) { const double a = A.x[d][i], b = A.x[d][j]; A.x[d][i] = b; A.x[d][j] = a; }
                    const int pa = A.perm[i], pb = A.perm[j];
                    A.perm[i] = pb; A.perm[j] = pa;
                }
            }
        }
        if (tid == 0) {
            A.t_np0[t] = np0;
            nchild_nodes[n] = (np0 > maxleaf) + (len - np0 > maxleaf);
        }
        __syncthreads();                                         // masks and counters are reused by the block's next node
    }
}

// ---- particle routing by the rank kd-tree (prepare_body_inOrderOf_domain, 1_Indexing/src/domains.c:163-296) -----------
// The same level-synchronous partition with GIVEN split values: the reference's bksort_body_inplace is, for runs of three
// or more, the standard pairing (k-th big element from the left of [0, ns) <-> k-th small one from the right of [ns, len),
// ns = number of elements <= split; model-checked in tests/test_device_tree_model.py); runs of one and two keep its
// special cases (two elements are put in ascending order whatever the split).
__global__ void route_flag_kernel(BuildArrays A, long long npart, int dir) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npart) return;
    const int t = A.seg[i];
    unsigned char f = 0;
    if (t >= 0 && A.t_len[t] > 2) f = A.x[dir][i] > A.t_split[t];
    A.flag[i] = f;
}

// per node: left count; runs of length <= 2 are also put in place here
__global__ void route_split_kernel(BuildArrays A, int lvl_begin, int lvl_count, int dir) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    const long long s = A.t_start[t];
    const int len = A.t_len[t];
    const double split = A.t_split[t];
    double* X = A.x[dir];
    int np0;
    if (len == 0) np0 = 0;
    else if (len == 1) np0 = X[s] > split ? 0 : 1;
    else if (len == 2) {
        if (X[s] > X[s + 1]) {
            for (int d = 0; d < 3; d++) { const double a = A.x[d][s]; A.x[d][s] = A.x[d][s + 1]; A.x[d][s + 1] = a; }
            const int p = A.perm[s]; A.perm[s] = A.perm[s + 1]; A.perm[s + 1] = p;
        }
        np0 = X[s] > split ? 0 : (X[s + 1] <= split ? 2 : 1);
    } else np0 = len - (int)(A.G[s + len] - A.G[s]);
    A.t_np0[t] = np0;
}

// both children of every node become nodes of the next level (possibly empty); their split values come from the rank tree
__global__ void route_children_kernel(BuildArrays A, int lvl_begin, int lvl_count, int next_begin, const double* __restrict__ dsplit,
                                      int next_heap_first, int nheap_inner) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    const int start = A.t_start[t], len = A.t_len[t], np0 = A.t_np0[t];
    for (int sde = 0; sde < 2; sde++) {
        const int c = next_begin + 2 * n + sde, h = next_heap_first + 2 * n + sde;
        A.t_start[c] = sde ? start + np0 : start;
        A.t_len[c] = sde ? len - np0 : np0;
        A.t_parent[c] = 2 * t + sde;
        A.t_split[c] = h < nheap_inner ? dsplit[h] : 0.0;
        A.t_child[2 * t + sde] = c;
    }
}

// ---- device-resident stepping: particles stay in HBM between steps -------------------------------
__global__ void iota_kernel(int* __restrict__ a, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}
// velocities and ids follow the permutation the tree build applied to the positions
__global__ void carry_kernel(const int* __restrict__ perm, long long n, const double* __restrict__ vx, const double* __restrict__ vy,
                             const double* __restrict__ vz, const int* __restrict__ gid, double* __restrict__ ox, double* __restrict__ oy,
                             double* __restrict__ oz, int* __restrict__ ogid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = perm[i];
    ox[i] = vx[p]; oy[i] = vy[p]; oz[i] = vz[p]; ogid[i] = gid[p];
}
// vel += acc * dkh (1_Indexing/src/photoNs.c:176-180; dkh = 0.5 dk G)
__global__ void kick_kernel(const float4* __restrict__ acc, long long n, double dkh, double* __restrict__ vx, double* __restrict__ vy,
                            double* __restrict__ vz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    vx[i] += (double)a.x * dkh; vy[i] += (double)a.y * dkh; vz[i] += (double)a.z * dkh;
}
// the same with the mid-field part of the short-range force (fp64, tree order) added: the reference kicks with the sum
// of P2P and M2L / L2L / L2P (1_Indexing/src/photoNs.c:161-180 after fmm_task / fmm_ext)
__global__ void kick_mid_kernel(const float4* __restrict__ acc, const double* __restrict__ mid, long long n, double dkh,
                                double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    vx[i] += ((double)a.x + mid[3 * i]) * dkh; vy[i] += ((double)a.y + mid[3 * i + 1]) * dkh; vz[i] += ((double)a.z + mid[3 * i + 2]) * dkh;
}
// pos += vel * dd, then wrapped into [0, box) exactly as the reference's while loops do (photoNs.c:182-208)
__global__ void drift_kernel(long long n, double dd, double box, const double* __restrict__ vx, const double* __restrict__ vy,
                             const double* __restrict__ vz, double* __restrict__ x, double* __restrict__ y, double* __restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {x[i] + vx[i] * dd, y[i] + vy[i] * dd, z[i] + vz[i] * dd};
    if (box > 0.0)
        for (int k = 0; k < 3; k++) {
            if (!(p[k] == p[k]) || fabs(p[k]) > 1e6 * box) continue;      // NaN / runaway: leave it, do not spin
            while (p[k] < 0.0) p[k] += box;
            while (p[k] >= box) p[k] -= box;
        }
    x[i] = p[0]; y[i] = p[1]; z[i] = p[2];
}
__global__ void aos_from_soa_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z, long long n,
                                    double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[3 * i] = x[i]; out[3 * i + 1] = y[i]; out[3 * i + 2] = z[i]; }
}

__global__ void widen_index_kernel(const int* __restrict__ perm, long long n, long long* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = perm[i];
}

__global__ void iota_offset_kernel(int* __restrict__ perm, long long n, int first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = first + (int)i;
}

// bottom-up: subtree leaf / node counts
__global__ void count_up_kernel(BuildArrays A, int lvl_begin, int lvl_count) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    int nl = 0, nn = 1;
    for (int s = 0; s < 2; s++) {
        const int c = A.t_child[2 * t + s];
        if (A.t_len[t] == 0) continue;             // an empty root has no children at all
        if (c >= 0) { nl += A.t_nleaf[c]; nn += A.t_nnode[c]; }
        else nl += 1;
    }
    A.t_nleaf[t] = nl;
    A.t_nnode[t] = nn;
}

struct TreeOut {
    int nleaf;               // offset of the nodes in the unified box array
    double* box;             // [(nleaf + nnode)][6] centre, width
    int* son;                // [nnode][2] unified ids (leaf l -> l, node n -> nleaf + n), -1 none
    int* node_npart;
    double* node_split;
    int* leaf_npart;
    int* leaf_ipart;
    int* max_width;                  // float bits (non-negative floats order like integers): [0] largest leaf width, [1] largest width of a node with a leaf child
    int* node_leaf0;                 // first leaf and number of leaves of every node's subtree (leaves are numbered left to right),
    int* node_nleaf;                 // for walks restricted to a range of target leaves
};

// top-down: final ids, kd cells, output arrays
__global__ void assign_down_kernel(BuildArrays A, TreeOut O, int lvl_begin, int lvl_count, int dir, double l0, double l1, double l2,
                                   double h0, double h1, double h2) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lvl_count) return;
    const int t = lvl_begin + n;
    double lo[3], hi[3];
    int id, leafbase;
    const int par = A.t_parent[t];
    if (par < 0) {
        lo[0] = l0; lo[1] = l1; lo[2] = l2; hi[0] = h0; hi[1] = h1; hi[2] = h2;
        id = 0;
        leafbase = 0;
    } else {
        const int p = par >> 1, side = par & 1;
        const int pd = (dir + 2) % 3;                         // the parent's split direction
        for (int k = 0; k < 3; k++) { lo[k] = A.t_lo[3 * p + k]; hi[k] = A.t_hi[3 * p + k]; }
        const double ps = A.t_split[p];
        if (side == 0) hi[pd] = ps; else lo[pd] = ps;
        const int c0 = A.t_child[2 * p];
        id = A.t_id[p] + 1;
        leafbase = A.t_leafbase[p];
        if (side == 1) {
            if (c0 >= 0) { id += A.t_nnode[c0]; leafbase += A.t_nleaf[c0]; }
            else leafbase += 1;
        }
    }
    for (int k = 0; k < 3; k++) { A.t_lo[3 * t + k] = lo[k]; A.t_hi[3 * t + k] = hi[k]; }
    A.t_id[t] = id;
    A.t_leafbase[t] = leafbase;
    double nwid[3], ncen[3];
    for (int k = 0; k < 3; k++) { nwid[k] = hi[k] - lo[k]; ncen[k] = 0.5 * (hi[k] + lo[k]); }
    double* nb = O.box + 6 * (size_t)(O.nleaf + id);
    for (int k = 0; k < 3; k++) { nb[k] = ncen[k]; nb[3 + k] = nwid[k]; }
    O.node_npart[id] = A.t_len[t];
    O.node_leaf0[id] = leafbase;
    O.node_nleaf[id] = A.t_nleaf[t];
    const double split = A.t_split[t];
    O.node_split[id] = split;
    if (A.t_len[t] == 0) { O.son[2 * id] = O.son[2 * id + 1] = -1; return; }
    int lb = leafbase, nid = id + 1, ip = A.t_start[t];
    double wmax = 0.0;
    for (int s = 0; s < 2; s++) {
        const int c = A.t_child[2 * t + s];
        if (c >= 0) {
            O.son[2 * id + s] = O.nleaf + nid;
            nid += A.t_nnode[c];
            lb += A.t_nleaf[c];
            ip += A.t_len[c];
        } else {
            const int cnt = -1 - c;
            O.son[2 * id + s] = lb;
            O.leaf_npart[lb] = cnt;
            O.leaf_ipart[lb] = ip;
            double* b = O.box + 6 * (size_t)lb;
            for (int k = 0; k < 3; k++) { b[k] = ncen[k]; b[3 + k] = nwid[k]; }
            if (s == 0) { b[3 + dir] = split - lo[dir]; b[dir] = 0.5 * (lo[dir] + split); }
            else        { b[3 + dir] = hi[dir] - split; b[dir] = 0.5 * (hi[dir] + split); }
            wmax = fmax(wmax, fmax(b[3], fmax(b[4], b[5])));
            lb += 1;
            ip += cnt;
        }
    }
    warp_atomic_max(O.max_width, __float_as_int((float)wmax));
    // largest cell of a node that has a leaf child: bounds every particle separation a LOCAL walk can list (a leaf pair enters
    // the frontier from a pair that passed the cutoff test with this node in the leaf's place)
    const double pmax = wmax > 0.0 ? fmax(nwid[0], fmax(nwid[1], nwid[2])) : 0.0;
    warp_atomic_max(O.max_width + 1, __float_as_int((float)pmax));
}

__global__ void soa_from_aos_kernel(const double* __restrict__ pos, long long n, double* __restrict__ x, double* __restrict__ y,
                                    double* __restrict__ z, int* __restrict__ perm, int* __restrict__ seg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = pos[3 * i];
    y[i] = pos[3 * i + 1];
    z[i] = pos[3 * i + 2];
    perm[i] = (int)i;
    seg[i] = 0;
}

__global__ void pack_fixed_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z, long long n,
                                  double ox, double oy, double oz, double scale, float mass, int4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 o;
    o.x = (int)(unsigned int)(__double2ll_rn((x[i] - ox) * scale) & 0xffffffffLL);     // same conversion as csr_pack.cuh:to_fixed
    o.y = (int)(unsigned int)(__double2ll_rn((y[i] - oy) * scale) & 0xffffffffLL);
    o.z = (int)(unsigned int)(__double2ll_rn((z[i] - oz) * scale) & 0xffffffffLL);
    o.w = __float_as_int(mass);
    out[i] = o;
}

__global__ void leaf_pack_kernel(const int* __restrict__ npart, const int* __restrict__ ipart, int n, int2* __restrict__ out,
                                 int* __restrict__ maxocc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_int2(ipart[i], npart[i]);
    warp_atomic_max(maxocc, npart[i]);
}

__global__ void acc_unpermute_kernel(const float4* __restrict__ acc, const int* __restrict__ perm, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    double* o = out + 3 * (size_t)perm[i];
    o[0] = a.x; o[1] = a.y; o[2] = a.z;
}

// ------------------------------------------------------------------------------------------------ walk
constexpr int kMaxPeers = 16;

struct WalkParams {
    const double* box;   // target (local) tree: unified [id][6] centre, width; leaves [0, nleaf), nodes nleaf + n
    const int* son;      // [node][2] unified ids
    int nleaf;
    double theta, rcut;
    double period;       // displacement unit of the images (BOXSIZE)
    double tc[3], tw[3]; // the local domain box the images / halos are pruned against (1_Indexing/src/remotes.c:374-386)
    // source trees: the local tree again (peer == me) and, in a multi-rank run, every other rank's tree
    const double* sbox;  // concatenated unified box arrays of all peers
    const int* sson;     // concatenated son arrays (peer-local unified ids)
    int me, npeer;
    long long sbox_base[kMaxPeers];   // first unified id of peer p in sbox
    long long sson_base[kMaxPeers];   // first node of peer p in sson
    int snleaf[kMaxPeers];
    int ts_base[kMaxPeers];           // task source id = ts_base[p] + leaf: 0 for me (local ids), ghost leaf ids otherwise
    const double* tb;    // tight bounds of the particles of every LOCAL leaf, [leaf][6] = lo[3], hi[3]; nullptr: no check
    const double* stb;   // the same for the source trees' leaves (concatenated per rank)
    long long stb_base[kMaxPeers];
    // M2L tasks (pairs the acceptance criterion hands to the multipole expansion), emitted when mt != nullptr:
    // target = local unified id, source = unified id inside rank `peer`, mq = (peer << 5) | displacement index
    int* mt;
    int* ms;
    int* mq;
    ull cap_m2l;
    // walk restricted to the target leaves [t_lo, t_hi) (t_hi <= 0: all): items whose target subtree misses the range are
    // dropped when they are generated, so a step can be split into target chunks whose lists are walked, packed and consumed
    // one after the other (bounded list memory); the union over a partition of the leaves is the full task multiset
    const int* node_leaf0;
    const int* node_nleaf;
    int t_lo, t_hi;
};

__device__ __forceinline__ bool target_in_range(const WalkParams& P, int id) {
    int l0 = id, n = 1;
    if (id >= P.nleaf) { l0 = __ldg(P.node_leaf0 + (id - P.nleaf)); n = __ldg(P.node_nleaf + (id - P.nleaf)); }
    return l0 < P.t_hi && l0 + n > P.t_lo;
}
// an M2L task whose target spans several chunks belongs to the chunk that holds the target's first leaf
__device__ __forceinline__ bool target_starts_in_range(const WalkParams& P, int id) {
    const int l0 = id >= P.nleaf ? __ldg(P.node_leaf0 + (id - P.nleaf)) : id;
    return l0 >= P.t_lo && l0 < P.t_hi;
}

__constant__ int c_shift[28][3];   // [27] = zero displacement walked with the remote rules (the reference's zero-shift self exchange)

// 1_Indexing/src/fmm.c:266-325
__device__ __forceinline__ int acceptance(const double wi[3], const double wj[3], const double dist[3], double theta, double rcut) {
    double w[3], mn[3];
    for (int k = 0; k < 3; k++) w[k] = (wi[k] + wj[k]) * 0.5;
    const double dd2 = dist[0] * dist[0] + dist[1] * dist[1] + dist[2] * dist[2];
    for (int k = 0; k < 3; k++) {
        double m = dist[k];
        if (m < 0.0) m = -m;
        m -= w[k];
        if (m <= 0.0) m = 0.0;
        mn[k] = m;
    }
    if (mn[0] + mn[1] + mn[2] < 0.0001) return 0;
    const double dm2 = mn[0] * mn[0] + mn[1] * mn[1] + mn[2] * mn[2];
    const double c2 = rcut * rcut;
    if (dm2 >= c2) return -1;
    if (dd2 > 1.0 * c2) return 0;
    double wmax = w[0];
    if (w[1] > wmax) wmax = w[1];
    if (w[2] > wmax) wmax = w[2];
    wmax *= 2;
    if (wmax * wmax < theta * theta * dd2) return 1;
    return 0;
}

// would prepare_sendtree2 have cut this node out of the image sent for displacement `disp`? (remotes.c:374-399)
__device__ __forceinline__ bool image_pruned(const WalkParams& P, const double nc[3], const double nw[3], const double disp[3]) {
    double dr = 0.0;
    for (int k = 0; k < 3; k++) {
        double d = P.tc[k] - nc[k] - disp[k];
        if (d < 0.0) d = -d;
        d -= (P.tw[k] + nw[k]) * 0.5;
        if (d > 0.0) dr += d * d;
    }
    dr = sqrt(dr);
    double wmax = nw[0];
    if (wmax < nw[1]) wmax = nw[1];
    if (wmax < nw[2]) wmax = nw[2];
    return dr >= P.rcut || wmax < 0.95 * P.theta * dr;
}

// item = target id (27 bits) | source id (27) | peer (4) | image (5)
__host__ __device__ __forceinline__ ull item(int im, int jm, int peer, int sh) {
    return ((ull)(unsigned)im << 36) | ((ull)(unsigned)jm << 9) | ((ull)(unsigned)peer << 5) | (ull)sh;
}

// One level of the breadth-first dual-tree walk.  Items are (target id, source id, source rank, image); an item
// either emits a leaf-leaf task, opens one side (2 items), opens both (self pair, 4 items) or dies.  The local
// tree against itself without displacement follows walk_task_p2p; every other (rank, image) combination follows
// walk_task_p2p_ext on the image prepare_sendtree2 would have sent, whose cut nodes are recognised on the fly.
//
// The levels run WITHOUT host round-trips: the frontier sizes live in a ring of three device counters (level L reads
// ring[L % 3], appends to ring[(L + 1) % 3] and clears ring[(L + 2) % 3] for the level after), the host enqueues a batch
// of levels back to back (an exhausted frontier makes the remaining launches no-ops) and reads the counters once.
// counters: [0..2] ring, [3] tasks, [4] M2L tasks, [5] minimal-image violations, [6] overflow flags, [7] items processed.
// Writes beyond a capacity are dropped and flagged (the host grows the buffers and repeats the walk).
constexpr int kWalkCounters = 8;
__global__ void __launch_bounds__(256, 4) walk_level_kernel(const ull* __restrict__ in, ull cap_in, ull* __restrict__ out, ull cap_out,
                                                            ull* __restrict__ counters, int* __restrict__ tt, int* __restrict__ ts,
                                                            ull cap_task, int level, WalkParams P) {
    __shared__ int s_cnt[3][8];
    __shared__ ull s_base[3];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const ull stride = (ull)gridDim.x * blockDim.x;
    ull n_in = counters[level % 3];
    if (n_in > cap_in) n_in = cap_in;                     // an overflowed level: what was dropped is flagged already
    // the slot the NEXT level appends to is cleared by every level, also by the no-op launches behind an exhausted frontier
    // (it still holds the input count of the level before this one)
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters[(level + 2) % 3] = 0; counters[7] += n_in; }
    if (n_in == 0) return;
    ull* const c_out = counters + (level + 1) % 3;
    const ull n_round = (n_in + 255) & ~255ull;           // whole blocks iterate together (block-wide slot claims)
    for (ull idx = (ull)blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += stride) {
        int nchild = 0;
        bool emit = false, m2l = false;
        int ci[4], cj[4];
        int im = 0, jm = 0, sh = 0, peer = 0, tsid = 0;
        if (idx < n_in) {
            const ull it = in[idx];
            im = (int)(it >> 36);
            jm = (int)((it >> 9) & 0x7ffffffu);
            peer = (int)((it >> 5) & 15);
            sh = (int)(it & 31);
            const int snl = P.snleaf[peer];
            const bool ileaf = im < P.nleaf, jleaf = jm < snl;
            const bool local = sh == 0 && peer == P.me;
            const int* __restrict__ sson = P.sson + 2 * P.sson_base[peer];
            tsid = P.ts_base[peer] + jm;
            if (local && im == jm) {
                if (ileaf) emit = true;
                else {
                    const int a0 = P.son[2 * (im - P.nleaf)], a1 = P.son[2 * (im - P.nleaf) + 1];
                    ci[0] = a0; cj[0] = a0; ci[1] = a0; cj[1] = a1; ci[2] = a1; cj[2] = a0; ci[3] = a1; cj[3] = a1;
                    nchild = 4;
                }
            } else if (ileaf && jleaf) {
                emit = true;
                if (P.period > 0.0 && P.tb) {
                    // the source will be read through minimal-image coordinates: every particle separation this leaf pair
                    // contains must stay below half the period, or another image would be picked.  tb: tight bounds of the
                    // leaves' particles {lo[3], hi[3]} (lo > hi for an empty leaf)
                    const double* bi = P.tb + 6 * (size_t)im;
                    const double* bj = P.stb + 6 * (size_t)(P.stb_base[peer] + jm);
                    bool bad = false;
                    if (bi[0] <= bi[3] && bj[0] <= bj[3])
                        for (int k = 0; k < 3; k++) {
                            const double disp = (double)c_shift[sh][k] * P.period;
                            const double d = fmax(bi[3 + k] - (bj[k] + disp), (bj[3 + k] + disp) - bi[k]);
                            bad |= !(d < 0.5 * P.period);
                        }
                    if (bad) atomicAdd(&counters[5], 1ull);
                }
            } else {
                const double* bi = P.box + 6 * (size_t)im;
                const double* bj = P.sbox + 6 * (size_t)(P.sbox_base[peer] + jm);
                double wi[3], wj[3], cjd[3], dist[3], disp[3];
                // read-only path (L1 / texture cache): siblings in the frontier share one of the two boxes
                const double2 i0 = __ldg(reinterpret_cast<const double2*>(bi)), i1 = __ldg(reinterpret_cast<const double2*>(bi) + 1),
                              i2 = __ldg(reinterpret_cast<const double2*>(bi) + 2);
                const double2 j0 = __ldg(reinterpret_cast<const double2*>(bj)), j1 = __ldg(reinterpret_cast<const double2*>(bj) + 1),
                              j2 = __ldg(reinterpret_cast<const double2*>(bj) + 2);
                const double ci_[3] = {i0.x, i0.y, i1.x}, cj_[3] = {j0.x, j0.y, j1.x};
                wi[0] = i1.y; wi[1] = i2.x; wi[2] = i2.y;
                wj[0] = j1.y; wj[1] = j2.x; wj[2] = j2.y;
                for (int k = 0; k < 3; k++) {
                    disp[k] = (double)c_shift[sh][k] * P.period;
                    cjd[k] = sh ? cj_[k] + disp[k] : cj_[k];
                    dist[k] = ci_[k] - cjd[k];
                }
                const int flag = acceptance(wi, wj, dist, P.theta, P.rcut);
                int open = 0;     // 1: target side, 2: source side
                if (local) {
                    if (flag == 0) {
                        if (ileaf) open = 2;
                        else if (jleaf) open = 1;
                        else open = (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) ? 1 : 2;
                    } else if (flag == 1) m2l = true;                       // walk_task_m2l, fmm.c:598,633,668
                } else if (flag != -1) {
                    bool pruned = false;
                    if (!jleaf) {
                        pruned = image_pruned(P, cj_, wj, disp);
                    }
                    if (ileaf) { if (flag != 1 && !pruned) open = 2; else m2l = true; }   // remotes.c:506: accepted OR cut by the sender
                    else if (jleaf) { if (flag != 1) open = 1; else m2l = true; }
                    else if (flag != 1) open = (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2] || pruned) ? 1 : 2;
                    else m2l = true;
                }
                if (open == 1) {
                    const int2 sn = __ldg(reinterpret_cast<const int2*>(P.son) + (im - P.nleaf));
                    ci[0] = sn.x; ci[1] = sn.y;
                    cj[0] = cj[1] = jm;
                    nchild = 2;
                } else if (open == 2) {
                    const int2 sn = __ldg(reinterpret_cast<const int2*>(sson) + (jm - snl));
                    cj[0] = sn.x; cj[1] = sn.y;
                    ci[0] = ci[1] = im;
                    nchild = 2;
                }
            }
        }
        if (P.t_hi > 0) {
            if (nchild && ci[0] != ci[nchild - 1]) {            // the target side was opened: keep the children that meet the range
                int m = 0;
                for (int k = 0; k < nchild; k++)
                    if (target_in_range(P, ci[k])) { ci[m] = ci[k]; cj[m] = cj[k]; m++; }
                nchild = m;
            }
            if (m2l) m2l = target_starts_in_range(P, im);
        }
        // claim output slots: per warp totals, ONE atomic per block and counter (all walkers of the chip add to the same three
        // words; per-warp atomics made the walk atomic-throughput bound), then per-warp and per-lane offsets
        int incl = nchild;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(full, incl, 31);
        const unsigned em = __ballot_sync(full, emit);
        const unsigned mm = P.mt ? __ballot_sync(full, m2l) : 0u;      // counted only when the M2L list is wanted
        if (lane == 0) { s_cnt[0][wid] = total; s_cnt[1][wid] = __popc(em); s_cnt[2][wid] = __popc(mm); }
        __syncthreads();
        if (threadIdx.x < 3) {
            int run = 0;
            for (int w = 0; w < 8; w++) { const int v = s_cnt[threadIdx.x][w]; s_cnt[threadIdx.x][w] = run; run += v; }
            ull* const ctr = threadIdx.x == 0 ? c_out : counters + 2 + threadIdx.x;          // frontier, tasks [3], M2L tasks [4]
            const ull cap = threadIdx.x == 0 ? cap_out : (threadIdx.x == 1 ? cap_task : P.cap_m2l);
            const ull b = run ? atomicAdd(ctr, (ull)run) : 0ull;
            if (run && b + (ull)run > cap) atomicOr(&counters[6], 1ull << threadIdx.x);
            s_base[threadIdx.x] = b;
        }
        __syncthreads();
        if (nchild) {
            const ull base = s_base[0] + (ull)s_cnt[0][wid] + (ull)(incl - nchild);
            if (base + nchild <= cap_out)
                for (int k = 0; k < nchild; k++) out[base + k] = item(ci[k], cj[k], peer, sh);
        }
        if (emit) {
            const ull tb = s_base[1] + (ull)s_cnt[1][wid] + (ull)__popc(em & ((1u << lane) - 1));
            if (tb < cap_task) { tt[tb] = im; ts[tb] = tsid; }
        }
        if (m2l && P.mt) {
            const ull mb = s_base[2] + (ull)s_cnt[2][wid] + (ull)__popc(mm & ((1u << lane) - 1));
            if (mb < P.cap_m2l) { P.mt[mb] = im; P.ms[mb] = jm; P.mq[mb] = (peer << 5) | sh; }
        }
        __syncthreads();
    }
}

// tight bounds of every local leaf's particles, from the fixed-point coordinates the force kernel reads
__global__ void leaf_bounds_kernel(const int2* __restrict__ leaf, int nleaf, const int4* __restrict__ part, double ox, double oy, double oz,
                                   double step, double* __restrict__ tb) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaf) return;
    const int2 L = leaf[l];
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0, 0, 0};
    for (int p = L.x; p < L.x + L.y; p++) {
        const int4 q = part[p];
        const unsigned v[3] = {(unsigned)q.x, (unsigned)q.y, (unsigned)q.z};
        for (int k = 0; k < 3; k++) { lo[k] = min(lo[k], v[k]); hi[k] = max(hi[k], v[k]); }
    }
    double* o = tb + 6 * (size_t)l;
    const double org[3] = {ox, oy, oz};
    for (int k = 0; k < 3; k++) {
        o[k] = L.y > 0 ? org[k] + (double)lo[k] * step : 1.0;
        o[3 + k] = L.y > 0 ? org[k] + (double)hi[k] * step : 0.0;
    }
}

// after sorting: a source listed twice in a row (the same leaf reached through two different images) cannot be
// represented with minimal-image sources
__global__ void csr_duplicate_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col, int nrow, unsigned int* __restrict__ dup) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    unsigned int d = 0;
    for (long long k = row_ptr[r] + 1; k < row_ptr[r + 1]; k++) d += col[k] == col[k - 1];
    if (d) atomicAdd(dup, d);
}

}  // namespace dt
}  // namespace p2p
