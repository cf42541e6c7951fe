// csr_pack.cuh -- interaction-list packing on the device.
//
// The reference re-packs and re-uploads per call: Indexing ships {t,s} pairs plus a padded
// [leaf][maxParts][3] fp64 position array (1_Indexing/src/fmm.c:851-877), Redundant ships a private
// copy of both leaves' particles per task (2_Redundant/src/fmm.c:812-838).  Here the walk's
// (target, source) pairs are turned ONCE per step into a CSR over target leaves:
//   count -> exclusive scan -> scatter -> per-row ascending sort of the source ids
// (the sort makes the row order, hence the FP32 summation order, independent of atomics timing and
// gives the staging copies ascending addresses).  All kernels are HBM-bound integer work.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace p2p {

// Counts tasks per target leaf.  Task ids are validated HERE (the host never walks the list): a task
// outside [row0, row0 + nrow) x [0,nsrc) raises *bad and is dropped by the scatter as well.  (row0, nrow) is the row
// window of the list: the whole leaf table, or the target range of one chunk of a chunked step -- every per-row pass of
// the packing then touches that window only.
__global__ void csr_count_kernel(const int* __restrict__ tt, const int* __restrict__ ts, long long n, int row0, int nrow, int nsrc,
                                 unsigned int* __restrict__ cnt, unsigned int* __restrict__ bad) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int t = tt[i], s = ts[i];
        if ((unsigned)(t - row0) < (unsigned)nrow && (unsigned)s < (unsigned)nsrc) atomicAdd(cnt + t, 1u);
        else atomicAdd(bad, 1u);
    }
}

// three-phase exclusive scan of unsigned counts into 64-bit offsets
constexpr int kScanTile = 2048;  // items per block (256 threads x 8)

__global__ void __launch_bounds__(256) scan_tile_sums_kernel(const unsigned int* __restrict__ cnt, int n,
                                                             unsigned long long* __restrict__ tile_sum) {
    __shared__ unsigned long long red[8];
    const int base = blockIdx.x * kScanTile;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int i = base + k * 256 + threadIdx.x;
        if (i < n) s += cnt[i];
    }
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += red[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(unsigned long long* __restrict__ tile_sum, int ntile) {
    // single block, serial over chunks of 1024 tiles (ntile <= a few thousand)
    __shared__ unsigned long long sh[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntile; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long v = i < ntile ? tile_sum[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            unsigned long long o = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
            __syncthreads();
            sh[threadIdx.x] += o;
            __syncthreads();
        }
        unsigned long long incl = sh[threadIdx.x];
        if (i < ntile) tile_sum[i] = carry + incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const unsigned int* __restrict__ cnt, int n,
                                                         const unsigned long long* __restrict__ tile_off,
                                                         long long* __restrict__ row_ptr,
                                                         unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long wsum[8];
    const int base = blockIdx.x * kScanTile;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // thread owns 8 consecutive items
    unsigned int v[8];
    unsigned long long s = 0;
    const int i0 = base + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = (i0 + k < n) ? cnt[i0 + k] : 0u; s += v[k]; }
    unsigned long long incl = s;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    unsigned long long off = tile_off[blockIdx.x];
    for (int k = 0; k < w; k++) off += wsum[k];
    unsigned long long run = off + incl - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (i0 + k < n) { row_ptr[i0 + k] = (long long)run; cursor[i0 + k] = run; }
        run += v[k];
    }
    if (i0 <= n - 1 && n - 1 < i0 + 8) row_ptr[n] = (long long)run;  // owner of the last item closes the array
}

// Tight bounds of every leaf's particles in the fixed-point frame: {centre xyz, -} and {half extent xyz, -}, both rounded
// outwards.  Differences are taken relative to the leaf's first particle, so a leaf that straddles the periodic wrap is
// handled like any other (a leaf spans far less than half the frame).
struct LeafBounds {
    int4 c, h;
};
__global__ void leaf_bounds_fixed_kernel(const int2* __restrict__ leaf, int nleaf, const int4* __restrict__ part, LeafBounds* __restrict__ out) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaf) return;
    const int2 L = leaf[l];
    LeafBounds b;
    b.c = make_int4(0, 0, 0, 0);
    b.h = make_int4(0, 0, 0, 0);
    if (L.y > 0) {
        const int4 r = part[L.x];
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int p = L.x + 1; p < L.x + L.y; p++) {
            const int4 q = part[p];
            const int d[3] = {q.x - r.x, q.y - r.y, q.z - r.z};          // wraps to the minimal image
            for (int k = 0; k < 3; k++) { lo[k] = min(lo[k], d[k]); hi[k] = max(hi[k], d[k]); }
        }
        const int r3[3] = {r.x, r.y, r.z};
        int c[3], h[3];
        for (int k = 0; k < 3; k++) {
            const long long mid = ((long long)lo[k] + hi[k]) >> 1;       // floor
            c[k] = (int)((unsigned)r3[k] + (unsigned)(int)mid);
            h[k] = (int)(((long long)hi[k] - lo[k] + 2) >> 1);           // covers both ends after the floor
        }
        b.c = make_int4(c[0], c[1], c[2], 0);
        b.h = make_int4(h[0], h[1], h[2], 0);
    }
    out[l] = b;
}
// squared gap between two leaves' bounds, in fixed-point steps (minimal image)
__device__ __forceinline__ double bounds_gap2(const LeafBounds& a, const LeafBounds& b) {
    const int d[3] = {b.c.x - a.c.x, b.c.y - a.c.y, b.c.z - a.c.z};
    const int h[3] = {a.h.x + b.h.x, a.h.y + b.h.y, a.h.z + b.h.z};
    double g2 = 0.0;
    for (int k = 0; k < 3; k++) {
        const long long g = llabs((long long)d[k]) - (long long)h[k];
        if (g > 0) g2 += (double)g * (double)g;
    }
    return g2;
}

// Scatter into the CSR rows.  With far2 > 0 (truncated kernel) every column is classified on the way: a source leaf whose
// bounds are at least sqrt(far2) fixed-point steps from the target leaf's is FAR (the force kernel's cheap far body is valid
// for every particle pair of such a leaf pair), every other column is NEAR and carries bit 31, so the unsigned row sort puts
// the far columns FIRST: the kernel sums the many small far terms before the few large near ones (FP32 rounding is
// relative to the running sum).
__global__ void csr_scatter_kernel(const int* __restrict__ tt, const int* __restrict__ ts, long long n, int row0, int nrow, int nsrc,
                                   unsigned long long* __restrict__ cursor, int* __restrict__ col, const LeafBounds* __restrict__ lb,
                                   double far2) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int t = tt[i], s = ts[i];
        if ((unsigned)(t - row0) >= (unsigned)nrow || (unsigned)s >= (unsigned)nsrc) continue;
        unsigned v = (unsigned)s | 0x80000000u;            // bit 31: NEAR class (every column when no classes are wanted)
        if (far2 > 0.0) {
            const LeafBounds a = lb[t], b = lb[s];
            if (bounds_gap2(a, b) >= far2) v = (unsigned)s;
        }
        unsigned long long p = atomicAdd(cursor + t, 1ull);
        col[p] = (int)v;
    }
}

// one warp per row: bitonic sort of the row's source ids (rows up to kSortCap)
constexpr int kSortCap = 2048;
constexpr int kSortRegCap = 512;     // rows up to here are sorted in registers

// Bitonic network over 32 E keys held E per lane, key i = e * 32 + lane (the layout of a coalesced load): exchanges at
// distance j < 32 are a shuffle, at distance j >= 32 they stay inside the lane.  No shared memory, no barriers: at the ~160
// columns of a 256^3 row the shared-memory form spent 3.2 ms per step on 1e8 columns, 2-way bank conflicts at the short
// distances included.
template <int E>
__device__ __forceinline__ void warp_sort_keys(unsigned int (&v)[E], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & je) == 0) {
                        const bool up = ((e << 5) & k) == 0;          // k >= 64 here: a bit of e
                        const unsigned int a = v[e], b = v[e | je];
                        v[e] = up ? min(a, b) : max(a, b);
                        v[e | je] = up ? max(a, b) : min(a, b);
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const unsigned int y = __shfl_xor_sync(0xffffffffu, v[e], j);
                    const bool up = (((e << 5) | lane) & k) == 0;
                    v[e] = (lower == up) ? min(v[e], y) : max(v[e], y);
                }
            }
        }
    }
}
template <int E>
__device__ __forceinline__ void warp_sort_row(unsigned int* __restrict__ col, long long b, int len, int lane) {
    unsigned int v[E];
#pragma unroll
    for (int e = 0; e < E; e++) v[e] = (e * 32 + lane < len) ? col[b + e * 32 + lane] : 0xffffffffu;
    warp_sort_keys<E>(v, lane);
#pragma unroll
    for (int e = 0; e < E; e++)
        if (e * 32 + lane < len) col[b + e * 32 + lane] = v[e];
}

// (rows above kSortCap columns are listed in long_rows[0 .. *unsorted) for csr_sort_long_rows_kernel)
__global__ void __launch_bounds__(128) csr_sort_rows_kernel(const long long* __restrict__ row_ptr, int nrow,
                                                            int* __restrict__ col_, unsigned int* __restrict__ unsorted,
                                                            int* __restrict__ long_rows) {
    __shared__ unsigned int sh[4][kSortCap];
    unsigned int* __restrict__ col = reinterpret_cast<unsigned int*>(col_);     // unsigned order: near columns (bit 31) last
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int* a = sh[w];
    for (int row = blockIdx.x * 4 + w; row < nrow; row += gridDim.x * 4) {
        const long long b = row_ptr[row];
        const int len = (int)(row_ptr[row + 1] - b);
        if (len <= 1) continue;
        if (len > kSortCap) { if (lane == 0) long_rows[atomicAdd(unsorted, 1u)] = row; continue; }
        if (len <= kSortRegCap) {
            if (len <= 32) warp_sort_row<1>(col, b, len, lane);
            else if (len <= 64) warp_sort_row<2>(col, b, len, lane);
            else if (len <= 128) warp_sort_row<4>(col, b, len, lane);
            else if (len <= 256) warp_sort_row<8>(col, b, len, lane);
            else warp_sort_row<16>(col, b, len, lane);
            continue;
        }
        int m = 1;
        while (m < len) m <<= 1;
        for (int i = lane; i < m; i += 32) a[i] = i < len ? col[b + i] : 0xffffffffu;
        __syncwarp();
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < (m >> 1); t += 32) {         // one compare-exchange per lane and step: no idle partner lanes
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                    const unsigned int x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
                __syncwarp();
            }
        }
        for (int i = lane; i < len; i += 32) col[b + i] = a[i];
        __syncwarp();
    }
}

// Rows above kSortCap columns (dense clumps): one block per row, bitonic network in global memory in the form whose
// compare-exchanges all put the smaller key at the lower index, so that positions beyond the row's length act as +inf
// without being stored.
__global__ void __launch_bounds__(256) csr_sort_long_rows_kernel(const long long* __restrict__ row_ptr, const unsigned int* __restrict__ n_long,
                                                                 const int* __restrict__ long_rows, int* __restrict__ col_) {
    unsigned int* __restrict__ col = reinterpret_cast<unsigned int*>(col_);
    const int n = (int)*n_long;                           // usually 0: scanning the whole row table for them cost 0.85 ms at 256^3
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int row = long_rows[r];
        const long long b = row_ptr[row];
        const long long len = row_ptr[row + 1] - b;
        long long m = 1;
        while (m < len) m <<= 1;
        for (long long k = 2; k <= m; k <<= 1) {
            for (long long j = k >> 1; j > 0; j >>= 1) {
                for (long long t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                    const long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const long long p = (j == (k >> 1)) ? (i ^ (k - 1)) : (i | j);      // first step of a merge: mirror partner
                    if (p < len) {
                        const unsigned int x = col[b + i], y = col[b + p];
                        if (x > y) { col[b + i] = y; col[b + p] = x; }
                    }
                }
                __syncthreads();
            }
        }
    }
}

// sum over tasks of n_t * n_s: the pair-interaction count of the metric; also records
// every row's work and a histogram of the target occupancy n_t for the row schedule
constexpr int kWorkBuckets = 64;
constexpr int kMinBandRows = 16384;  // smallest band of the row schedule (bounds the size of the band histograms)

// Band size of the row schedule, chosen on the device from the leaf occupancies: (number of distinct target
// occupancies) x (resident warps), so that the warps resident at any time work on one or two occupancy values
// (see below), but no larger, so that a band's particles stay in L2.  occ[0..63]: scratch, zeroed by the caller.
__global__ void __launch_bounds__(256) occupancy_hist_kernel(const int2* __restrict__ leaf, int nrow, unsigned int* __restrict__ occ) {
    __shared__ unsigned int h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < nrow) { const int nt = leaf[row].y; if (nt > 0) atomicAdd(h + min(nt, 63), 1u); }
    __syncthreads();
    if (threadIdx.x < 64 && h[threadIdx.x]) atomicAdd(occ + threadIdx.x, h[threadIdx.x]);
}
__global__ void band_rows_kernel(const unsigned int* __restrict__ occ, int nrow, int resident_warps, int fixed, int* __restrict__ band_rows) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int distinct = 0;
        for (int b = 0; b < 64; b++) distinct += occ[b] != 0;
        const int target = fixed > 0 ? max(fixed, kMinBandRows) : max(kMinBandRows, distinct * resident_warps);
        const int nband = max(1, (nrow + target / 2) / target);          // equal bands: no short last band
        *band_rows = max(kMinBandRows, (nrow + nband - 1) / nband);
    }
}
// (row_ptr, row_work, row_mid are the caller's arrays advanced to the first row of the window; leaf is the whole table)
__global__ void __launch_bounds__(256) pair_count_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col,
                                                         const int2* __restrict__ leaf, int row0, int nrow,
                                                         unsigned long long* __restrict__ npairs,
                                                         unsigned long long* __restrict__ row_work,
                                                         unsigned int* __restrict__ hist, const int* __restrict__ band_rows_p,
                                                         int* __restrict__ row_mid) {
    const int band_rows = *band_rows_p;
    unsigned long long s = 0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int row = warp; row < nrow; row += nwarp) {
        const long long b = row_ptr[row], e = row_ptr[row + 1];
        const unsigned long long nt = (unsigned long long)leaf[row0 + row].y;
        unsigned long long ns = 0;
        int nfar = 0;                                    // columns without the near bit: the row's leading part after the sort
        for (long long i = b + lane; i < e; i += 32) {
            const unsigned int c = (unsigned int)col[i];
            ns += (unsigned long long)leaf[c & 0x7fffffffu].y;
            nfar += (c >> 31) == 0;
        }
        for (int d = 16; d >= 1; d >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, d); nfar += __shfl_xor_sync(0xffffffffu, nfar, d); }
        const unsigned long long w = nt * ns;
        if (lane == 0) {
            row_mid[row] = nfar;
            row_work[row] = w;
            if (w) atomicAdd(hist + (row / band_rows) * kWorkBuckets + min((int)nt, kWorkBuckets - 1), 1u);
            s += w;
        }
    }
    if (lane == 0 && s) atomicAdd(npairs, s);
}

// Row schedule: only rows that have work; bands of consecutive rows (a spatially compact piece of the kd order whose
// particles stay in L2; size from band_rows_kernel), and inside a band by target occupancy n_t (fullest leaves first).  The force kernel instantiates its slice code per number of target pairs; with rows
// in arbitrary order the 16 resident warps of an SM run up to 8 different ~7 KB code bodies at once and the
// instruction cache thrashes (ncu on the clustered box: stall_no_instruction 9.5 per issue, issue rate
// halved).  Grouping rows by n_t makes the whole chip run the same one or two bodies at any time; the
// fullest (most expensive) rows go first, so the tail of the persistent kernel consists of cheap rows.
// Ordering the WHOLE list by n_t instead swept the particle array once per occupancy value: 8.5 GB of DRAM reads per
// launch at 256^3 against 1 GB of compulsory traffic (ncu, profiles/r1e_ncu_rows_kernel_256_final.txt).
__global__ void __launch_bounds__(1024) work_bucket_offsets_kernel(const unsigned int* __restrict__ hist, unsigned int* __restrict__ cursor,
                                                                  int nband, unsigned int* __restrict__ n_active) {
    // nband: upper bound (rows / kMinBandRows + 1); bands beyond the real count are empty.  Exclusive scan of the histogram
    // in schedule order (band major, fullest occupancy first) by one block.
    __shared__ unsigned int ws[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int n = nband * kWorkBuckets;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int slot = i < n ? (i / kWorkBuckets) * kWorkBuckets + (kWorkBuckets - 1 - i % kWorkBuckets) : 0;
        const unsigned int v = i < n ? hist[slot] : 0;
        unsigned int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            const unsigned int w = ws[threadIdx.x];
            unsigned int z = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, z, o);
                if (threadIdx.x >= o) z += y;
            }
            ws[threadIdx.x] = z - w;
        }
        __syncthreads();
        const unsigned int excl = carry + ws[threadIdx.x >> 5] + x - v;
        if (i < n) cursor[slot] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_active = carry;
}
__global__ void work_order_scatter_kernel(const unsigned long long* __restrict__ row_work, const int2* __restrict__ nt_of,
                                          int row0, int nrow, unsigned int* __restrict__ cursor, int* __restrict__ order, const int* __restrict__ band_rows_p) {
    const int band_rows = *band_rows_p;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < nrow) {
        const unsigned long long w = row_work[row];
        if (w) order[atomicAdd(cursor + (row / band_rows) * kWorkBuckets + min(nt_of[row0 + row].y, kWorkBuckets - 1), 1u)] = row0 + row;
    }
}

__global__ void add_counter_kernel(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst += *src;
}

// 32-bit fixed-point coordinate over the (padded) box: q = round((x - origin) * 2^32 / extent) mod 2^32.
// Differences of two such coordinates wrap to the minimal image, so displaced (periodic image /
// halo) copies of a particle map to the same value as the original.
__device__ __forceinline__ int to_fixed(double x, double origin, double inv_step) {
    const long long q = __double2ll_rn((x - origin) * inv_step);
    return (int)(unsigned int)(q & 0xffffffffLL);
}
// fp64 xyz rows (staged in device memory) -> int4 {xi, yi, zi, mass bits}
__global__ void pack_particles_kernel(const double* __restrict__ xyz, long long n, double ox, double oy, double oz,
                                      double inv_step, float mass, int4* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_int4(to_fixed(xyz[3 * i], ox, inv_step), to_fixed(xyz[3 * i + 1], oy, inv_step),
                                  to_fixed(xyz[3 * i + 2], oz, inv_step), __float_as_int(mass));
}
// float4 {x, y, z, m} already on the device (NCCL halo buffers) -> int4
__global__ void pack_particles_f4_kernel(const float4* __restrict__ in, long long n, double ox, double oy, double oz,
                                         double inv_step, int4* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 p = in[i];
        out[i] = make_int4(to_fixed((double)p.x, ox, inv_step), to_fixed((double)p.y, oy, inv_step),
                           to_fixed((double)p.z, oz, inv_step), __float_as_int(p.w));
    }
}
__global__ void leaves_pack_kernel(const int* __restrict__ start, const int* __restrict__ count, int n, int start_off,
                                   int2* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_int2(start[i] + start_off, count[i]);
}
__global__ void add_offset_kernel(int* __restrict__ v, long long n, int off) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) v[i] += off;
}
__global__ void deinterleave_kernel(const int* __restrict__ pairs, long long n, int off, int* __restrict__ tt,
                                    int* __restrict__ ts) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { tt[i] = pairs[2 * i]; ts[i] = pairs[2 * i + 1] + off; }
}
__global__ void acc_to_f64_kernel(const float4* __restrict__ acc, long long n, double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { float4 a = acc[i]; out[3 * i] = a.x; out[3 * i + 1] = a.y; out[3 * i + 2] = a.z; }
}

}  // namespace p2p
