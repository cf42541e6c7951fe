// p2p_api.cu -- native C-ABI (include/p2p_b200.h) over the sm_100a kernels.
// Host-side state is one context per device; device buffers are grow-only and persist across steps
// (the reference allocates once, sized from the first step, and never re-sizes: SURVEY defect D16).
// There is NO CPU fallback: without a device every call that needs one returns P2P_ERR_NODEVICE.
#include "../../include/p2p_b200.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "csr_pack.cuh"
#include "halo.cuh"
#include "p2p_ctx.h"
#include "p2p_gcoef.h"
#include "p2p_kernel.cuh"

namespace {
thread_local char g_err[512] = "";
int g_verbose = -1;
}  // namespace

int p2p_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (g_verbose < 0) g_verbose = getenv("P2P_B200_VERBOSE") ? 1 : 0;
    if (g_verbose) fprintf(stderr, "[p2p_b200] error %d: %s\n", code, g_err);
    return code;
}
#define fail p2p_fail

int p2p_update_occupancy(p2p_ctx* c) {
    CU(cudaMemsetAsync(c->d_occ, 0, 64 * sizeof(unsigned int), c->stream));
    if (c->nleaf) {
        p2p::occupancy_hist_kernel<<<(c->nleaf + 255) / 256, 256, 0, c->stream>>>(c->leaf.p, c->nleaf, c->d_occ);
        CU(cudaGetLastError());
    }
    return 0;
}

int p2p_use(p2p_ctx* c) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    return 0;
}

namespace {

// bounding cube of host positions, doubled, so that no separation reaches extent / 2
int auto_box(p2p_ctx* c, const double* pos, long long stride, long long n) {
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    bool first = true;
    for (long long i = 0; i < n; i++) {
        const double* p = pos + i * stride;
        if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;   // padding slots of the compat layouts
        for (int k = 0; k < 3; k++) {
            if (first || p[k] < lo[k]) lo[k] = p[k];
            if (first || p[k] > hi[k]) hi[k] = p[k];
        }
        first = false;
    }
    double w = 0.0;
    for (int k = 0; k < 3; k++) w = std::max(w, hi[k] - lo[k]);
    if (!(w > 0.0)) w = 1.0;
    c->extent = 4.0 * w;
    for (int k = 0; k < 3; k++) c->origin[k] = 0.5 * (lo[k] + hi[k]) - 0.5 * c->extent;
    return 0;
}

int pinned(p2p_ctx* c, size_t bytes, void** out) {
    if (bytes > c->h_pinned_bytes) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        c->h_pinned = nullptr;
        c->h_pinned_bytes = 0;
        CU(cudaMallocHost(&c->h_pinned, bytes));
        c->h_pinned_bytes = bytes;
    }
    *out = c->h_pinned;
    return 0;
}

// host rows of 3 doubles (stride in doubles) -> device fixed-point int4 at dst[0..n)
int upload_xyz(p2p_ctx* c, const double* pos, long long stride, long long n, int4* dst) {
    if (n == 0) return 0;
    CU(c->stage.reserve((size_t)n * 24, c->stream));
    if (stride == 3) {
        CU(cudaMemcpyAsync(c->stage.p, pos, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
    } else {
        CU(cudaMemcpy2DAsync(c->stage.p, 24, pos, (size_t)stride * 8, 24, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    const int B = 256;
    p2p::pack_particles_kernel<<<(unsigned)((n + B - 1) / B), B, 0, c->stream>>>(
        reinterpret_cast<const double*>(c->stage.p), n, c->origin[0], c->origin[1], c->origin[2], 4294967296.0 / c->extent,
        (float)c->mass, dst);
    CU(cudaGetLastError());
    return 0;
}

int upload_ints(p2p_ctx* c, const int* a, const int* b, long long n, int** da, int** db) {
    // two int arrays through one staging buffer
    CU(c->itmp.reserve((size_t)(2 * n + 2), c->stream));
    *da = c->itmp.p;
    *db = c->itmp.p + n;
    if (n) {
        CU(cudaMemcpyAsync(*da, a, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(*db, b, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

constexpr int kStage = 384;   // particles per staging buffer: 2 x 6 KB + targets per warp -> 16 warps / SM fit
// defaults of the second-generation kernel (sweeps: profiles/r2_sweep_*.txt): one source per lane (half the code size of two:
// the clustered box runs many different row occupancies at once and the instruction cache holds them), 3 blocks / SM
constexpr int kDefaultNsrc = 1, kDefaultMinBlocks = 3;

template <int TT, int NSRC, bool TRUNC, bool PACKED, int MINB, int POLY, int STAGE = kStage>
int launch_rows(p2p_ctx* c, const p2p::KernelParams& P) {
    auto kern = p2p::p2p_rows_kernel<TT, NSRC, STAGE, TRUNC, PACKED, MINB, POLY>;
    const int smem = 4 * (int)sizeof(p2p::WarpSmem<TT, STAGE>);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) return fail(P2P_ERR_CUDA, "force kernel does not fit on an SM (smem %d)", smem);
    long long want = ((long long)P.nrow + 3) / 4;
    int grid = (int)std::min<long long>((long long)c->num_sm * per_sm, std::max<long long>(want, 1));
    kern<<<grid, 128, smem, c->stream>>>(P);
    CU(cudaGetLastError());
    c->last_blocks_per_sm = per_sm;
    return 0;
}

// First-generation kernel (16 targets per pass, one slice body): kept as the scalar cross-check variant and as the
// A/B baseline of the sweeps (tools/sweep.py), in its final round-1 tuning only.
template <int TT, bool TRUNC, bool PACKED>
int launch_cfg(p2p_ctx* c, const p2p::KernelParams& P) {
    return launch_rows<TT, 2, TRUNC, PACKED, 4, (TRUNC && PACKED) ? 1 : 0>(c, P);
}

// Second-generation kernel (p2p_rows2_kernel): one pass per row, near and far slice bodies.
template <int NSRC, bool TRUNC, int MINB, int DBG, bool BLOCKED, bool RETIRE>
int launch_rows2_inst(p2p_ctx* c, const p2p::KernelParams& P) {
    auto kern = p2p::p2p_rows2_kernel<NSRC, kStage, TRUNC, MINB, DBG, BLOCKED, RETIRE>;
    const int smem = 4 * (int)sizeof(p2p::WarpSmem2<kStage>);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) return fail(P2P_ERR_CUDA, "force kernel does not fit on an SM (smem %d)", smem);
    long long want = ((long long)P.nrow + 3) / 4;
    long long grid = std::min<long long>((long long)c->num_sm * per_sm, std::max<long long>(want, 1));
    p2p::KernelParams Q = P;
    Q.persist_blocks = (int)grid;
    // non-persistent mode: budgeted blocks in front of one persistent wave; a budgeted block handles at least four rows, blocks
    // that find the schedule exhausted leave at once
    if (RETIRE && P.rows_per_warp > 0) grid += want / 2;
    kern<<<(unsigned)grid, 128, smem, c->stream>>>(Q);
    CU(cudaGetLastError());
    c->last_blocks_per_sm = per_sm;
    return 0;
}
// the shipped configuration comes in four builds (blocked summation x retiring warps, both compile-time); the tuning and
// debug variants only in the general one
template <int NSRC, bool TRUNC, int MINB, int DBG = 0, bool ALL = false>
int launch_rows2(p2p_ctx* c, const p2p::KernelParams& P, bool blocked) {
    const bool retire = P.rows_per_warp > 0;
    if constexpr (ALL) {
        if (!blocked && !retire) return launch_rows2_inst<NSRC, TRUNC, MINB, DBG, false, false>(c, P);
        if (!blocked && retire) return launch_rows2_inst<NSRC, TRUNC, MINB, DBG, false, true>(c, P);
        if (blocked && !retire) return launch_rows2_inst<NSRC, TRUNC, MINB, DBG, true, false>(c, P);
    }
    p2p::KernelParams Q = P;
    if (!blocked) Q.block_leaves = 0x3fffffff;       // the general build: one block per class
    return launch_rows2_inst<NSRC, TRUNC, MINB, DBG, true, true>(c, Q);
}
template <bool TRUNC>
int launch_cfg2(p2p_ctx* c, const p2p::KernelParams& P, int nsrc, int minb, bool blocked) {
    if constexpr (TRUNC) {
        if (nsrc == 9) return launch_rows2<1, true, 2, 1>(c, P, blocked);          // error-budget variant: fp64 force factor (pair_exact2)
        if (nsrc == 2) return launch_rows2<2, true, 3>(c, P, blocked);
        return minb == 4 ? launch_rows2<1, true, 4>(c, P, blocked) : launch_rows2<1, true, 3, 0, true>(c, P, blocked);
    } else {
        return launch_rows2<1, false, 3>(c, P, blocked);
    }
}

}  // namespace

extern "C" {

const char* p2p_last_error(void) { return g_err; }

int p2p_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int p2p_create(p2p_ctx** out, int device) {
    if (!out) return fail(P2P_ERR_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(P2P_ERR_NODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(P2P_ERR_ARG, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(P2P_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    p2p_ctx* c = new p2p_ctx();
    c->device = device;
    c->num_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaMalloc(&c->d_counter, 4 * sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_npairs, sizeof(unsigned long long)));
    CU(cudaMemset(c->d_npairs, 0, sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_npairs_acc, sizeof(unsigned long long)));
    CU(cudaMemset(c->d_npairs_acc, 0, sizeof(unsigned long long)));
    CU(cudaMallocHost(&c->h_flags, 8 * sizeof(unsigned int)));
    memset(c->h_flags, 0, 8 * sizeof(unsigned int));
    CU(cudaMalloc(&c->d_counter2, 4 * sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_bad, 2 * sizeof(unsigned int)));       // [0] tasks with ids out of range, [1] oversize source leaf met by the force kernel
    CU(cudaMalloc(&c->d_occ, 64 * sizeof(unsigned int)));
    CU(cudaMemset(c->d_occ, 0, 64 * sizeof(unsigned int)));
    CU(cudaMemset(c->d_bad, 0, 2 * sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_npairs2, sizeof(unsigned long long)));
    {   // the copy / packing stream outranks the force kernel's: its small kernels take the room left for them first
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi));
    }
    for (int k = 0; k < 2; k++) {
        CU(cudaEventCreateWithFlags(&c->ev_packed[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_bounds, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_bounds, c->own_stream));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreate(&c->ev2));
    CU(cudaEventCreate(&c->ev3));
    CU(cudaEventCreate(&c->ev0_b));
    CU(cudaEventCreate(&c->ev1_b));
    CU(cudaEventCreate(&c->ev2_b));
    CU(cudaEventCreate(&c->ev3_b));
    CU(cudaMallocHost(&c->h_halo, 64 * sizeof(long long)));
    CU(cudaMalloc(&c->d_halo, 64 * sizeof(long long)));
    *out = c;
    return 0;
}

int p2p_destroy(p2p_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->part.release(); c->acc.release(); c->leaf.release(); c->tt.release(); c->ts.release(); c->col.release();
    c->itmp.release(); c->row_ptr.release(); c->cnt.release(); c->cursor.release(); c->tile.release(); c->stage.release();
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->d_npairs) cudaFree(c->d_npairs);
    if (c->d_npairs_acc) cudaFree(c->d_npairs_acc);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    c->acc64.release();
    c->row_work.release(); c->row_work2.release(); c->order.release(); c->order2.release(); c->row_mid.release(); c->row_mid2.release(); c->lbounds.release(); c->whist.release(); c->whist2.release();
    c->tt2.release(); c->ts2.release(); c->col2.release(); c->row_ptr2.release(); c->cnt2.release(); c->cursor2.release(); c->tile2.release();
    if (c->d_counter2) cudaFree(c->d_counter2);
    if (c->d_bad) cudaFree(c->d_bad);
    if (c->d_occ) cudaFree(c->d_occ);
    if (c->d_npairs2) cudaFree(c->d_npairs2);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int k = 0; k < 2; k++) { if (c->ev_packed[k]) cudaEventDestroy(c->ev_packed[k]); if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]); }
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_bounds) cudaEventDestroy(c->ev_bounds);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->ev3);
    cudaEventDestroy(c->ev0_b); cudaEventDestroy(c->ev1_b); cudaEventDestroy(c->ev2_b); cudaEventDestroy(c->ev3_b);
    if (c->h_halo) cudaFreeHost(c->h_halo);
    if (c->d_halo) cudaFree(c->d_halo);
    c->halo_cnt.release(); c->halo_off.release(); c->halo_tile.release(); c->halo_cursor.release();
    if (c->dtree) p2p_dtree_release(c->dtree);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int p2p_set_physics(p2p_ctx* c, double mass, double eps, double rs) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    if (!(eps >= 0.0) || !isfinite(mass)) return fail(P2P_ERR_ARG, "bad physics (mass %g eps %g)", mass, eps);
    c->mass = mass; c->eps = eps; c->rs = rs;
    return 0;
}

int p2p_set_box(p2p_ctx* c, const double origin[3], double extent) {
    if (!c || !origin || !(extent > 0.0)) return fail(P2P_ERR_ARG, "bad box");
    for (int k = 0; k < 3; k++) c->origin[k] = origin[k];
    c->extent = extent;
    c->box_set = true;
    return 0;
}

int p2p_set_tuning(p2p_ctx* c, int tt, int nsrc, int minb) {
    // tt 0 / 32: second-generation kernel (nsrc 1 / 2, min_blocks 3 / 4); tt 8 / 16: first-generation kernel in its final tuning
    if (!c || (tt && tt != 8 && tt != 16 && tt != 32) || (nsrc != 0 && nsrc != 1 && nsrc != 2 && nsrc != 9) || (minb != 0 && minb != 3 && minb != 4))
        return fail(P2P_ERR_ARG, "bad tuning (targets_per_pass 0/32 or 8/16, sources_per_lane 1/2, min_blocks 3/4)");
    c->tune_tt = tt; c->tune_nsrc = nsrc; c->tune_minb = minb;
    return 0;
}

int p2p_set_far_threshold(p2p_ctx* c, double u_far) {
    if (!c || !(u_far <= 0.0 || u_far >= P2P_U_FAR - 1e-12)) return fail(P2P_ERR_ARG, "far threshold must be 0 (off), < 0 (default) or >= %g", (double)P2P_U_FAR);
    c->far_u = u_far;
    c->csr_valid = false;
    return 0;
}

int p2p_set_kernel_variant(p2p_ctx* c, int v) {
    if (!c || v < 0 || v > 2) return fail(P2P_ERR_ARG, "bad kernel variant");
    c->variant = v;
    return 0;
}

int p2p_set_stream(p2p_ctx* c, void* s) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

int p2p_swap_lists(p2p_ctx* c) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    std::swap(c->tt, c->tt2); std::swap(c->ts, c->ts2); std::swap(c->col, c->col2); std::swap(c->row_ptr, c->row_ptr2);
    std::swap(c->cnt, c->cnt2); std::swap(c->cursor, c->cursor2); std::swap(c->tile, c->tile2); std::swap(c->row_work, c->row_work2);
    std::swap(c->order, c->order2); std::swap(c->row_mid, c->row_mid2); std::swap(c->whist, c->whist2);
    std::swap(c->d_counter, c->d_counter2); std::swap(c->d_npairs, c->d_npairs2);
    std::swap(c->ntask, c->ntask_b); std::swap(c->npairs, c->npairs_b); std::swap(c->csr_valid, c->csr_valid_b);
    std::swap(c->ev0, c->ev0_b); std::swap(c->ev1, c->ev1_b); std::swap(c->ev2, c->ev2_b); std::swap(c->ev3, c->ev3_b);
    std::swap(c->ms_compute, c->ms_compute_b); std::swap(c->ms_csr, c->ms_csr_b);
    std::swap(c->timed_compute, c->timed_compute_b); std::swap(c->timed_csr, c->timed_csr_b);
    return 0;
}

int p2p_set_force_blocks(p2p_ctx* c, int rows_per_warp) {
    if (!c || rows_per_warp < 0) return fail(P2P_ERR_ARG, "bad rows_per_warp");
    c->rows_per_warp = rows_per_warp;
    return 0;
}

int p2p_reserve_ghosts(p2p_ctx* c, int nghostleaf, int64_t nghost) {
    USE(c);
    if (nghostleaf < 0 || nghost < 0) return fail(P2P_ERR_ARG, "bad ghost capacities");
    CU(c->part.reserve((size_t)(c->npart + nghost) + 1, c->stream, (size_t)c->npart));
    CU(c->leaf.reserve((size_t)c->nleaf + nghostleaf + 1, c->stream, (size_t)c->nleaf));
    CU(c->lbounds.reserve(2 * ((size_t)c->nleaf + nghostleaf) + 2, c->stream, 2 * (size_t)std::min(c->bounds_n, c->nleaf)));
    return 0;
}

int p2p_upload_particles(p2p_ctx* c, const double* pos, int64_t stride, int64_t npart) {
    USE(c);
    if (npart < 0 || (npart && !pos) || stride < 3) return fail(P2P_ERR_ARG, "bad particle array");
    c->npart = npart; c->nghost = 0; c->nghostleaf = 0; c->csr_valid = false; c->bounds_n = 0;
    CU(c->part.reserve((size_t)npart + 1, c->stream));
    CU(c->acc.reserve((size_t)npart + 1, c->stream));
    if (!c->box_set) auto_box(c, pos, stride, npart);
    int r = upload_xyz(c, pos, stride, npart, c->part.p);
    if (r) return r;
    CU(cudaMemsetAsync(c->acc.p, 0, (size_t)npart * sizeof(float4), c->stream));
    CU(cudaMemsetAsync(c->d_npairs_acc, 0, sizeof(unsigned long long), c->stream));
    c->acc_tasks = 0;
    return 0;
}

int p2p_upload_leaves(p2p_ctx* c, const int* leaf_npart, const int* leaf_ipart, int nleaf) {
    USE(c);
    if (nleaf < 0 || (nleaf && (!leaf_npart || !leaf_ipart))) return fail(P2P_ERR_ARG, "bad leaf arrays");
    int mx = 0;
    for (int i = 0; i < nleaf; i++) {
        if (leaf_npart[i] < 0 || leaf_ipart[i] < 0 || (long long)leaf_ipart[i] + leaf_npart[i] > c->npart)
            return fail(P2P_ERR_ARG, "leaf %d {npart %d, ipart %d} outside the %lld uploaded particles", i, leaf_npart[i],
                        leaf_ipart[i], c->npart);
        mx = std::max(mx, leaf_npart[i]);
    }
    if (mx > P2P_MAX_LEAF) return fail(P2P_ERR_ARG, "leaf occupancy %d exceeds P2P_MAX_LEAF %d", mx, P2P_MAX_LEAF);
    c->nleaf = nleaf; c->nghostleaf = 0; c->nghost = 0; c->max_target_leaf = mx; c->csr_valid = false; c->bounds_n = 0;
    CU(c->leaf.reserve((size_t)nleaf + 1, c->stream));
    int *dc, *ds;
    int r = upload_ints(c, leaf_npart, leaf_ipart, nleaf, &dc, &ds);
    if (r) return r;
    if (nleaf) {
        p2p::leaves_pack_kernel<<<(nleaf + 255) / 256, 256, 0, c->stream>>>(ds, dc, nleaf, 0, c->leaf.p);
        CU(cudaGetLastError());
    }
    return p2p_update_occupancy(c);
}

static int append_ghost_leaves(p2p_ctx* c, const int* start, const int* count, int nleaf, long long nbody, int* first_id) {
    for (int i = 0; i < nleaf; i++) {
        if (count[i] < 0 || start[i] < 0 || (long long)start[i] + count[i] > nbody)
            return fail(P2P_ERR_ARG, "ghost leaf %d {start %d, count %d} outside the batch of %lld bodies", i, start[i],
                        count[i], nbody);
        if (count[i] > kStage) return fail(P2P_ERR_ARG, "ghost leaf %d too large (%d > %d)", i, count[i], kStage);
    }
    const int first = c->nleaf + c->nghostleaf;
    CU(c->leaf.reserve((size_t)first + nleaf + 1, c->stream, (size_t)first));
    int *ds, *dc;
    int r = upload_ints(c, start, count, nleaf, &ds, &dc);
    if (r) return r;
    if (nleaf) {
        p2p::leaves_pack_kernel<<<(nleaf + 255) / 256, 256, 0, c->stream>>>(ds, dc, nleaf, (int)(c->npart + c->nghost),
                                                                             c->leaf.p + first);
        CU(cudaGetLastError());
    }
    if (first_id) *first_id = first;
    c->nghostleaf += nleaf;
    c->nghost += nbody;
    c->csr_valid = false; c->bounds_n = std::min(c->bounds_n, first);
    return 0;
}

int p2p_append_ghosts(p2p_ctx* c, const double* pos, int64_t stride, int64_t nbody, const int* start, const int* count,
                      int nleaf, int* first_leaf_id) {
    USE(c);
    if (nbody < 0 || nleaf < 0 || (nbody && !pos) || (nleaf && (!start || !count)) || stride < 3)
        return fail(P2P_ERR_ARG, "bad ghost batch");
    if (!(c->extent > 0.0)) return fail(P2P_ERR_STATE, "upload local particles (or set the box) before ghosts");
    const long long base = c->npart + c->nghost;
    CU(c->part.reserve((size_t)(base + nbody) + 1, c->stream, (size_t)base));
    int r = upload_xyz(c, pos, stride, nbody, c->part.p + base);
    if (r) return r;
    return append_ghost_leaves(c, start, count, nleaf, nbody, first_leaf_id);
}

int p2p_append_ghosts_device(p2p_ctx* c, const void* d_xyzm, int64_t nbody, const int* start, const int* count, int nleaf,
                             int* first_leaf_id) {
    USE(c);
    if (nbody < 0 || nleaf < 0 || (nbody && !d_xyzm) || (nleaf && (!start || !count))) return fail(P2P_ERR_ARG, "bad ghost batch");
    if (!(c->extent > 0.0)) return fail(P2P_ERR_STATE, "upload local particles (or set the box) before ghosts");
    const long long base = c->npart + c->nghost;
    CU(c->part.reserve((size_t)(base + nbody) + 1, c->stream, (size_t)base));
    if (nbody) {
        p2p::pack_particles_f4_kernel<<<(unsigned)((nbody + 255) / 256), 256, 0, c->stream>>>(
            reinterpret_cast<const float4*>(d_xyzm), nbody, c->origin[0], c->origin[1], c->origin[2], 4294967296.0 / c->extent,
            c->part.p + base);
        CU(cudaGetLastError());
    }
    return append_ghost_leaves(c, start, count, nleaf, nbody, first_leaf_id);
}

int p2p_clear_ghosts(p2p_ctx* c) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    c->nghost = 0; c->nghostleaf = 0; c->csr_valid = false; c->bounds_n = std::min(c->bounds_n, c->nleaf);
    return 0;
}

int p2p_clear_tasks(p2p_ctx* c) {
    if (!c) return fail(P2P_ERR_ARG, "null context");
    c->ntask = 0; c->csr_valid = false; c->npairs = -1;
    return 0;
}

int p2p_append_tasks(p2p_ctx* c, const int* tt, const int* ts, int64_t n, int off) {
    USE(c);
    if (n < 0 || (n && (!tt || !ts))) return fail(P2P_ERR_ARG, "bad task arrays");
    CU(c->tt.reserve((size_t)(c->ntask + n) + 1, c->stream, (size_t)c->ntask));
    CU(c->ts.reserve((size_t)(c->ntask + n) + 1, c->stream, (size_t)c->ntask));
    if (n) {
        CU(cudaMemcpyAsync(c->tt.p + c->ntask, tt, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->ts.p + c->ntask, ts, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        if (off) {
            p2p::add_offset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->ts.p + c->ntask, n, off);
            CU(cudaGetLastError());
        }
    }
    c->ntask += n; c->csr_valid = false;
    return 0;
}

int p2p_append_tasks_interleaved(p2p_ctx* c, const int* pairs, int64_t n, int off) {
    USE(c);
    if (n < 0 || (n && !pairs)) return fail(P2P_ERR_ARG, "bad task array");
    CU(c->tt.reserve((size_t)(c->ntask + n) + 1, c->stream, (size_t)c->ntask));
    CU(c->ts.reserve((size_t)(c->ntask + n) + 1, c->stream, (size_t)c->ntask));
    if (n) {
        CU(c->itmp.reserve((size_t)(2 * n), c->stream));
        CU(cudaMemcpyAsync(c->itmp.p, pairs, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        p2p::deinterleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->itmp.p, n, off, c->tt.p + c->ntask,
                                                                                    c->ts.p + c->ntask);
        CU(cudaGetLastError());
    }
    c->ntask += n; c->csr_valid = false;
    return 0;
}

namespace {
struct ListSet {                       // one set of task / CSR buffers (the context owns two)
    DevBuf<int>*tt, *ts, *col;
    DevBuf<long long>* row_ptr;
    DevBuf<unsigned int>* cnt;
    DevBuf<unsigned long long>*cursor, *tile, *row_work;
    DevBuf<int>*order, *row_mid;
    DevBuf<unsigned int>* whist;
    unsigned int* d_counter;           // [0] row scheduler, [1] unsorted rows, [2] rows in the work-ordered schedule
    unsigned long long* d_npairs;
};
// rows per band of the row schedule (csr_pack.cuh): chosen on the device unless P2P_B200_BAND_ROWS fixes it (sweeps)
int band_rows() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("P2P_B200_BAND_ROWS");
        v = e ? atoi(e) : -1;
        if (v < 256) v = -1;                       // -1: chosen on the device (band_rows_kernel)
    }
    return v;
}
ListSet list_set(p2p_ctx* c, int k) {
    if (k == 0)
        return ListSet{&c->tt, &c->ts, &c->col, &c->row_ptr, &c->cnt, &c->cursor, &c->tile, &c->row_work, &c->order, &c->row_mid, &c->whist,
                       c->d_counter, c->d_npairs};
    return ListSet{&c->tt2, &c->ts2, &c->col2, &c->row_ptr2, &c->cnt2, &c->cursor2, &c->tile2, &c->row_work2, &c->order2, &c->row_mid2,
                   &c->whist2, c->d_counter2, c->d_npairs2};
}

int reserve_csr(p2p_ctx* c, const ListSet& L, long long n, cudaStream_t st) {
    const int nrow = c->nleaf;
    CU(L.row_ptr->reserve((size_t)nrow + 2, st));
    CU(L.cnt->reserve((size_t)nrow + 1, st));
    CU(L.cursor->reserve((size_t)nrow + 1, st));
    CU(L.col->reserve((size_t)n + 1, st));
    CU(L.tile->reserve((size_t)((nrow + p2p::kScanTile - 1) / p2p::kScanTile) + 1, st));
    CU(L.row_work->reserve((size_t)nrow + 1, st));
    CU(L.order->reserve((size_t)nrow + 1, st));
    CU(L.row_mid->reserve((size_t)nrow + 1, st));
    CU(L.whist->reserve(2 * (size_t)p2p::kWorkBuckets * (nrow / p2p::kMinBandRows + 1) + 128, st));
    return 0;
}

// squared near / far threshold in fixed-point steps (0: the kernel has no far class -- plain Newtonian kernel, first-
// generation or scalar variant)
double far_threshold2(const p2p_ctx* c) {
    const bool v2 = c->variant != P2P_KERNEL_SCALAR && (c->tune_tt == 0 || c->tune_tt == 32);
    if (!(c->rs > 0.0) || !v2 || c->far_u == 0.0 || !(c->extent > 0.0)) return 0.0;
    const double u = c->far_u > 0.0 ? c->far_u : (double)P2P_U_FAR;
    const double d = u * 2.0 * c->rs / (c->extent / 4294967296.0);
    return d * d;
}

// tight fixed-point bounds of all leaves (local + ghost), recomputed when particles or leaves changed
int leaf_bounds_fixed(p2p_ctx* c, cudaStream_t st) {
    const int n = c->nleaf + c->nghostleaf, have = c->bounds_n;
    if (have >= n) return 0;
    // only the leaves that are new (the ghost leaves of a halo that arrived after the local list was packed): the entries
    // of the leaves before them may be in use by a packing kernel on another stream
    CU(c->lbounds.reserve(2 * (size_t)n + 2, st, 2 * (size_t)have));
    p2p::leaf_bounds_fixed_kernel<<<(n - have + 127) / 128, 128, 0, st>>>(c->leaf.p + have, n - have, c->part.p,
                                                                             reinterpret_cast<p2p::LeafBounds*>(c->lbounds.p) + have);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_bounds, st));
    c->bounds_n = n;
    return 0;
}

// count -> scan -> scatter (+ near / far classification) -> sort -> pair count of the n tasks in L.tt / L.ts, all on stream st
int pack_csr(p2p_ctx* c, const ListSet& L, long long n, cudaStream_t st) {
    const int nleaf = c->nleaf;
    int r = reserve_csr(c, L, n, st);
    if (r) return r;
    // row window: the whole leaf table, or the target range of the chunk being packed (p2p_forces_local); rows outside the
    // window keep whatever an earlier packing left and are never scheduled
    const int row0 = c->row_hi > c->row_lo ? c->row_lo : 0;
    const int nrow = c->row_hi > c->row_lo ? std::min(c->row_hi, nleaf) - row0 : nleaf;
    const double far2 = far_threshold2(c);
    const bool v2 = c->variant != P2P_KERNEL_SCALAR && (c->tune_tt == 0 || c->tune_tt == 32);
    if (far2 > 0.0 || v2) {                                // near / far classes, and the row reference points of the force kernel
        if ((r = leaf_bounds_fixed(c, st))) return r;
        CU(cudaStreamWaitEvent(st, c->ev_bounds, 0));      // bounds of earlier leaves may have been computed on another stream
    }
    const int ntile = (nrow + p2p::kScanTile - 1) / p2p::kScanTile;
    CU(cudaMemsetAsync(L.cnt->p + row0, 0, ((size_t)nrow + 1) * 4, st));
    CU(cudaMemsetAsync(L.d_counter, 0, 4 * sizeof(unsigned int), st));
    const int nband = nrow / p2p::kMinBandRows + 1;                    // upper bound; the band size itself is chosen on the device
    CU(cudaMemsetAsync(L.whist->p, 0, (2 * (size_t)p2p::kWorkBuckets * nband + 128) * sizeof(unsigned int), st));
    CU(cudaMemsetAsync(L.d_npairs, 0, sizeof(unsigned long long), st));
    if (nrow <= 0) {
        CU(cudaMemsetAsync(L.row_ptr->p, 0, sizeof(long long), st));
        if (n > 0) CU(cudaMemsetAsync(c->d_bad, 0xff, sizeof(unsigned int), st));          // every task is out of range
        return 0;
    }
    const int G = c->num_sm * 8;
    const int nsrc = c->nleaf + c->nghostleaf;
    if (n) {
        p2p::csr_count_kernel<<<G, 256, 0, st>>>(L.tt->p, L.ts->p, n, row0, nrow, nsrc, L.cnt->p, c->d_bad);
        CU(cudaGetLastError());
    }
    p2p::scan_tile_sums_kernel<<<ntile, 256, 0, st>>>(L.cnt->p + row0, nrow, L.tile->p);
    p2p::scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(L.tile->p, ntile);
    p2p::scan_apply_kernel<<<ntile, 256, 0, st>>>(L.cnt->p + row0, nrow, L.tile->p, L.row_ptr->p + row0, L.cursor->p + row0);
    CU(cudaGetLastError());
    if (n) {
        p2p::csr_scatter_kernel<<<G, 256, 0, st>>>(L.tt->p, L.ts->p, n, row0, nrow, nsrc, L.cursor->p, L.col->p,
                                                   reinterpret_cast<const p2p::LeafBounds*>(c->lbounds.p), far2);
        // (the list of long rows borrows the schedule array, which is written only afterwards)
        p2p::csr_sort_rows_kernel<<<G, 128, 0, st>>>(L.row_ptr->p + row0, nrow, L.col->p, L.d_counter + 1, L.order->p);
        p2p::csr_sort_long_rows_kernel<<<c->num_sm * 2, 256, 0, st>>>(L.row_ptr->p + row0, L.d_counter + 1, L.order->p, L.col->p);
        int* d_band = reinterpret_cast<int*>(L.whist->p + 2 * (size_t)p2p::kWorkBuckets * nband);
        p2p::band_rows_kernel<<<1, 32, 0, st>>>(c->d_occ, nrow, c->num_sm * 16, band_rows(), d_band);
        p2p::pair_count_kernel<<<G, 256, 0, st>>>(L.row_ptr->p + row0, L.col->p, c->leaf.p, row0, nrow, L.d_npairs, L.row_work->p + row0, L.whist->p,
                                                  d_band, L.row_mid->p + row0);
        p2p::work_bucket_offsets_kernel<<<1, 1024, 0, st>>>(L.whist->p, L.whist->p + p2p::kWorkBuckets * nband, nband, L.d_counter + 2);
        p2p::work_order_scatter_kernel<<<(nrow + 255) / 256, 256, 0, st>>>(L.row_work->p + row0, c->leaf.p, row0, nrow,
                                                                          L.whist->p + p2p::kWorkBuckets * nband, L.order->p, d_band);
        CU(cudaGetLastError());
    }
    return 0;
}
}  // namespace

int p2p_build_csr(p2p_ctx* c) {
    USE(c);
    CU(cudaEventRecord(c->ev2, c->stream));
    const ListSet L = list_set(c, 0);
    // (only the list flag: a force kernel consuming the other list set may be raising the kernel flag d_bad[1] right now)
    CU(cudaMemsetAsync(c->d_bad, 0, sizeof(unsigned int), c->stream));
    int r = pack_csr(c, L, c->ntask, c->stream);
    if (r) return r;
    CU(cudaMemcpyAsync(c->h_flags + 2, c->d_bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev3, c->stream));
    c->timed_csr = true;
    c->csr_valid = true;
    c->flags_pending = true;
    c->npairs = -1;
    return 0;
}

// The list is validated on the device while it is counted; the verdict is read at the next point
// that synchronises anyway (download / counts / synchronize).
static int check_flags(p2p_ctx* c) {
    if (!c->flags_pending) return 0;
    c->flags_pending = false;
    if (c->h_flags[2] != 0) {
        c->csr_valid = false;
        return fail(P2P_ERR_ARG, "%u task(s) reference leaves outside [0,%d) x [0,%d)", c->h_flags[2], c->nleaf,
                    c->nleaf + c->nghostleaf);
    }
    if (c->h_flags[3] != 0) {
        c->h_flags[3] = 0;
        cudaMemsetAsync(c->d_bad + 1, 0, sizeof(unsigned int), c->stream);       // reported once
        return fail(P2P_ERR_ARG, "a source leaf holds more than %d particles (ghost leaf table not validated?): its pairs were skipped", kStage);
    }
    return 0;
}

namespace {
// launches the force kernel over the CSR of L on stream st and adds its pair count to the running total
int launch_force(p2p_ctx* c, const ListSet& L, long long ntask, cudaStream_t st) {
    if (c->max_target_leaf > P2P_MAX_LEAF) return fail(P2P_ERR_ARG, "target leaves above %d particles are not supported", P2P_MAX_LEAF);
    p2p::KernelParams P;
    memset(&P, 0, sizeof P);
    P.part = c->part.p; P.leaf = c->leaf.p; P.row_ptr = L.row_ptr->p; P.col = L.col->p; P.acc = c->acc.p;
    P.counter = L.d_counter; P.n_active = L.d_counter + 2; P.row_order = L.order->p; P.nrow = c->nleaf;
    P.row_mid = L.row_mid->p; P.err = c->d_bad + 1; P.rows_per_warp = c->rows_per_warp;
    // blocked summation of long classes: off unless P2P_B200_BLOCK_LEAVES names a block size (experiments)
    static int blk = -1;
    if (blk < 0) { const char* e = getenv("P2P_B200_BLOCK_LEAVES"); blk = e ? atoi(e) : 0; if (blk < 0) blk = 0; }
    P.block_leaves = blk > 0 ? blk : p2p::kBlockLeaves;
    const bool blocked = blk > 0;
    const bool trunc = c->rs > 0.0;
    const bool packed = c->variant != P2P_KERNEL_SCALAR;
    const bool v2 = packed && (c->tune_tt == 0 || c->tune_tt == 32);
    // kernel length unit.  Truncated kernel: L* = 2 r_s / sqrt(log2 e) makes exp(-u^2) = 2^(-r'^2); the second-generation
    // kernel takes the power-of-two fraction of the frame nearest to it, L0 = extent / 2^p, so that the fixed-point step is
    // 2^(p - 32) units EXACTLY, and carries kappa = (L0 / L*)^2 in its coefficients.  Plain kernel: the frame itself.
    const double sl2e = sqrt(1.4426950408889634);
    double unit = trunc ? 2.0 * c->rs / sl2e : (c->extent > 0.0 ? c->extent : 1.0);
    double kappa = 1.0;
    if (trunc && v2 && c->extent > 0.0) {
        const double lstar = unit;
        unit = c->extent * exp2(-round(log2(c->extent / lstar)));
        kappa = (unit / lstar) * (unit / lstar);
    }
    P.k_fix = (float)(c->extent / 4294967296.0 / unit);
    P.eps2 = (float)((c->eps / unit) * (c->eps / unit));
    P.neps2 = -P.eps2;
    P.nkappa = (float)-kappa;
    P.u_scale = (float)(sqrt(kappa) / sl2e);
    if (trunc) {
        // rinv' Q(u) = rinv' + v (c0 + c1 v + ...), v = r / L*: c_j = q_{j+2} / sl2e^(j+2); in units of L0: c_j kappa^(1 + j/2)
        for (int j = 0; j < p2p::kPolyTerms; j++) P.c[j] = (float)(P2P_GCOEF_10[j + 2] / pow(sl2e, j + 2) * pow(kappa, 1.0 + 0.5 * j));
        static_assert(p2p::kFarDegree == P2P_FAR_DEGREE, "far-field polynomial degree");
        // far field: kappa^(3/2) 2^(-kappa w') H(t), t = t' / kappa, t' = 1 / (w' + shift / kappa)
        for (int j = 0; j < p2p::kFarTerms; j++) P.cf[j] = (float)(P2P_GFAR[j] * pow(kappa, 1.5 - j));
        P.far_s0 = (float)P2P_FAR_SHIFT;
        P.far_shift = (float)(P2P_FAR_SHIFT / kappa);
        // dummy (padding) source: 2^(-r^2) flushes to exactly 0 there while r^8 still fits FP32, whatever the extent of the target leaf
        P.far_coord = 1000.0f * (float)sl2e;
    } else {
        P.far_coord = 1.0e18f;
    }
    P.lbounds = (v2 && c->bounds_n >= c->nleaf && c->nleaf > 0) ? c->lbounds.p : nullptr;
    P.out_scale = (float)(c->mass / (unit * unit));
    CU(cudaMemsetAsync(L.d_counter, 0, sizeof(unsigned int), st));
    int r = 0;
    if (c->nleaf > 0 && ntask > 0) {
        cudaStream_t keep = c->stream;
        c->stream = st;                                         // the launchers use c->stream
        if (v2) {
            const int nsrc = c->tune_nsrc ? c->tune_nsrc : kDefaultNsrc;
            const int minb = c->tune_minb ? c->tune_minb : kDefaultMinBlocks;
            r = trunc ? launch_cfg2<true>(c, P, nsrc, minb, blocked) : launch_cfg2<false>(c, P, nsrc, minb, blocked);
        } else {
            const int tt = (c->tune_tt == 8 || c->tune_tt == 16) ? c->tune_tt : (c->max_target_leaf <= 8 ? 8 : 16);
            if (tt == 8) {
                if (trunc) r = packed ? launch_cfg<8, true, true>(c, P) : launch_cfg<8, true, false>(c, P);
                else r = packed ? launch_cfg<8, false, true>(c, P) : launch_cfg<8, false, false>(c, P);
            } else {
                if (trunc) r = packed ? launch_cfg<16, true, true>(c, P) : launch_cfg<16, true, false>(c, P);
                else r = packed ? launch_cfg<16, false, true>(c, P) : launch_cfg<16, false, false>(c, P);
            }
        }
        c->stream = keep;
    }
    if (r) return r;
    p2p::add_counter_kernel<<<1, 32, 0, st>>>(L.d_npairs, c->d_npairs_acc);   // no host sync
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_flags + 3, c->d_bad + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    c->flags_pending = true;
    c->acc_tasks += ntask;
    return 0;
}
}  // namespace

int p2p_compute(p2p_ctx* c) {
    USE(c);
    if (!c->csr_valid) return fail(P2P_ERR_STATE, "p2p_compute before p2p_build_csr");
    CU(cudaEventRecord(c->ev0, c->stream));
    int r = launch_force(c, list_set(c, 0), c->ntask, c->stream);
    if (r) return r;
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed_compute = true;
    return 0;
}

int p2p_accumulated_counts(p2p_ctx* c, int64_t* ntask, int64_t* npairs) {
    USE(c);
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->d_npairs_acc, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int r = check_flags(c);
    if (r) return r;
    if (ntask) *ntask = c->acc_tasks;
    if (npairs) *npairs = (long long)v;
    return 0;
}

int p2p_zero_acc(p2p_ctx* c) {
    USE(c);
    if (c->npart) CU(cudaMemsetAsync(c->acc.p, 0, (size_t)c->npart * sizeof(float4), c->stream));
    CU(cudaMemsetAsync(c->d_npairs_acc, 0, sizeof(unsigned long long), c->stream));
    c->acc_tasks = 0;
    return 0;
}

int p2p_synchronize(p2p_ctx* c) {
    USE(c);
    CU(cudaStreamSynchronize(c->stream));
    return check_flags(c);
}

int p2p_download_acc(p2p_ctx* c, double* acc, int64_t stride, int accumulate) {
    USE(c);
    if ((c->npart && !acc) || stride < 3) return fail(P2P_ERR_ARG, "bad acc array");
    const long long n = c->npart;
    if (n == 0) { CU(cudaStreamSynchronize(c->stream)); return check_flags(c); }
    if (stride == 3 && !accumulate) {
        // packed destination: convert on the device and copy straight into the caller's buffer
        // (full PCIe rate when that buffer is pinned)
        CU(c->acc64.reserve((size_t)n * 3, c->stream));
        p2p::acc_to_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->acc.p, n, c->acc64.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(acc, c->acc64.p, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return check_flags(c);
    }
    void* hp;
    int r = pinned(c, (size_t)n * sizeof(float4), &hp);
    if (r) return r;
    CU(cudaMemcpyAsync(hp, c->acc.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((r = check_flags(c))) return r;
    const float4* a = reinterpret_cast<const float4*>(hp);
    if (accumulate) {
        for (long long i = 0; i < n; i++) { double* d = acc + i * stride; d[0] += a[i].x; d[1] += a[i].y; d[2] += a[i].z; }
    } else {
        for (long long i = 0; i < n; i++) { double* d = acc + i * stride; d[0] = a[i].x; d[1] = a[i].y; d[2] = a[i].z; }
    }
    return 0;
}

int p2p_counts(p2p_ctx* c, int64_t* ntask, int64_t* npairs) {
    USE(c);
    if (ntask) *ntask = c->ntask;
    if (npairs) {
        if (!c->csr_valid) return fail(P2P_ERR_STATE, "pair count needs p2p_build_csr");
        if (c->npairs < 0) {
            unsigned long long v = 0;
            CU(cudaMemcpyAsync(&v, c->d_npairs, sizeof v, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            int r = check_flags(c);
            if (r) return r;
            c->npairs = (long long)v;
        }
        *npairs = c->npairs;
    }
    return 0;
}

int p2p_download_csr(p2p_ctx* c, int64_t* row_ptr, int* col) {
    USE(c);
    if (!c->csr_valid) return fail(P2P_ERR_STATE, "no CSR built");
    if (row_ptr) CU(cudaMemcpyAsync(row_ptr, c->row_ptr.p, ((size_t)c->nleaf + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (col && c->ntask) CU(cudaMemcpyAsync(col, c->col.p, (size_t)c->ntask * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (col) for (long long i = 0; i < c->ntask; i++) col[i] &= 0x7fffffff;       // bit 31: class of the column
    return 0;
}

int p2p_download_csr_class(p2p_ctx* c, unsigned char* is_far, int* row_far) {
    USE(c);
    if (!c->csr_valid) return fail(P2P_ERR_STATE, "no CSR built");
    if (is_far && c->ntask) {
        int* tmp = (int*)malloc((size_t)c->ntask * 4);
        if (!tmp) return fail(P2P_ERR_ARG, "out of host memory");
        cudaError_t e = cudaMemcpyAsync(tmp, c->col.p, (size_t)c->ntask * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) for (long long i = 0; i < c->ntask; i++) is_far[i] = (unsigned char)(((unsigned)tmp[i] >> 31) ^ 1u);
        free(tmp);
        CU(e);
    }
    if (row_far && c->nleaf) {
        if (c->ntask) CU(cudaMemcpyAsync(row_far, c->row_mid.p, (size_t)c->nleaf * 4, cudaMemcpyDeviceToHost, c->stream));
        else memset(row_far, 0, (size_t)c->nleaf * 4);
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int p2p_last_timings(p2p_ctx* c, float* ms_compute, float* ms_csr) {
    USE(c);
    CU(cudaStreamSynchronize(c->stream));
    if (c->timed_compute) CU(cudaEventElapsedTime(&c->ms_compute, c->ev0, c->ev1));
    if (c->timed_csr) CU(cudaEventElapsedTime(&c->ms_csr, c->ev2, c->ev3));
    if (ms_compute) *ms_compute = c->ms_compute;
    if (ms_csr) *ms_csr = c->ms_csr;
    return 0;
}

int p2p_step_host(p2p_ctx* c, const double* pos, int64_t pos_stride, int64_t npart, const int* leaf_npart,
                  const int* leaf_ipart, int nleaf, const int* tt, const int* ts, int64_t ntask, double* acc,
                  int64_t acc_stride, int accumulate) {
    int r;
    if ((r = p2p_upload_particles(c, pos, pos_stride, npart))) return r;
    if ((r = p2p_upload_leaves(c, leaf_npart, leaf_ipart, nleaf))) return r;
    if ((r = p2p_clear_tasks(c))) return r;
    if ((r = p2p_append_tasks(c, tt, ts, ntask, 0))) return r;
    if ((r = p2p_build_csr(c))) return r;
    if ((r = p2p_compute(c))) return r;
    return p2p_download_acc(c, acc, acc_stride, accumulate);
}

int p2p_step_host_chunked(p2p_ctx* c, const double* pos, int64_t pos_stride, int64_t npart, const int* leaf_npart,
                          const int* leaf_ipart, int nleaf, const double* ghost_pos, int64_t ghost_stride, int64_t nghost,
                          const int* ghost_start, const int* ghost_count, int nghostleaf, const int* tt, const int* ts,
                          const int64_t* chunk_off, int nchunk, double* acc, int64_t acc_stride, int accumulate) {
    USE(c);
    if (nchunk < 0 || (nchunk && (!chunk_off || !tt || !ts))) return fail(P2P_ERR_ARG, "bad chunk arrays");
    int r;
    if ((r = p2p_upload_particles(c, pos, pos_stride, npart))) return r;
    if ((r = p2p_upload_leaves(c, leaf_npart, leaf_ipart, nleaf))) return r;
    if (nghostleaf > 0 || nghost > 0) {
        int first = 0;
        if ((r = p2p_append_ghosts(c, ghost_pos, ghost_stride, nghost, ghost_start, ghost_count, nghostleaf, &first))) return r;
    }
    cudaStream_t S0 = c->stream, S1 = c->copy_stream;
    long long maxn = 0;
    for (int g = 0; g < nchunk; g++) {
        if (chunk_off[g + 1] < chunk_off[g]) return fail(P2P_ERR_ARG, "chunk offsets must be non-decreasing");
        maxn = std::max<long long>(maxn, chunk_off[g + 1] - chunk_off[g]);
    }
    // both buffer sets are sized up front: nothing may be reallocated while the pipeline is in flight
    for (int k = 0; k < 2; k++) {
        const ListSet L = list_set(c, k);
        CU(L.tt->reserve((size_t)maxn + 1, S0));
        CU(L.ts->reserve((size_t)maxn + 1, S0));
        if ((r = reserve_csr(c, L, maxn, S0))) return r;
    }
    CU(cudaMemsetAsync(c->d_bad, 0, 2 * sizeof(unsigned int), S0));
    CU(cudaEventRecord(c->ev_ready, S0));                    // particles, leaves and ghosts are queued on S0
    CU(cudaStreamWaitEvent(S1, c->ev_ready, 0));
    CU(cudaEventRecord(c->ev0, S0));
    for (int g = 0; g < nchunk; g++) {
        const int b = g & 1;
        const ListSet L = list_set(c, b);
        const long long n = chunk_off[g + 1] - chunk_off[g];
        if (g >= 2) CU(cudaStreamWaitEvent(S1, c->ev_done[b], 0));      // the kernel of chunk g-2 has released set b
        if (n) {
            CU(cudaMemcpyAsync(L.tt->p, tt + chunk_off[g], (size_t)n * 4, cudaMemcpyHostToDevice, S1));
            CU(cudaMemcpyAsync(L.ts->p, ts + chunk_off[g], (size_t)n * 4, cudaMemcpyHostToDevice, S1));
        }
        if ((r = pack_csr(c, L, n, S1))) return r;
        CU(cudaEventRecord(c->ev_packed[b], S1));
        CU(cudaStreamWaitEvent(S0, c->ev_packed[b], 0));                 // copy + packing of chunk g overlapped kernel g-1
        if ((r = launch_force(c, L, n, S0))) return r;
        CU(cudaEventRecord(c->ev_done[b], S0));
    }
    CU(cudaEventRecord(c->ev1, S0));
    c->timed_compute = true;
    c->csr_valid = false;                                   // the resident CSR is only the last chunk's
    c->ntask = 0;
    CU(cudaMemcpyAsync(c->h_flags + 2, c->d_bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, S0));
    c->flags_pending = true;
    return p2p_download_acc(c, acc, acc_stride, accumulate);
}

// ---- halo planning of the multi-rank device path (halo.cuh) ------------------------------------------------------------
namespace {
// exclusive scan of c->halo_cnt[0, n) into c->halo_off[0, n]
int halo_scan(p2p_ctx* c, long long n, cudaStream_t st) {
    if (n > 0x7fffffffLL) return fail(P2P_ERR_ARG, "halo plan too large");
    const int ntile = (int)((n + p2p::kScanTile - 1) / p2p::kScanTile);
    CU(c->halo_off.reserve((size_t)n + 2, st));
    CU(c->halo_cursor.reserve((size_t)n + 2, st));
    CU(c->halo_tile.reserve((size_t)ntile + 2, st));
    if (n == 0) { CU(cudaMemsetAsync(c->halo_off.p, 0, sizeof(long long), st)); return 0; }
    p2p::scan_tile_sums_kernel<<<ntile, 256, 0, st>>>(c->halo_cnt.p, (int)n, c->halo_tile.p);
    p2p::scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(c->halo_tile.p, ntile);
    p2p::scan_apply_kernel<<<ntile, 256, 0, st>>>(c->halo_cnt.p, (int)n, c->halo_tile.p, c->halo_off.p, c->halo_cursor.p);
    CU(cudaGetLastError());
    return 0;
}
// totals of the nseg segments [bound[s], bound[s + 1]) of the scan -> host (synchronises the stream)
int halo_totals(p2p_ctx* c, const int* bound, int nseg, long long* out, cudaStream_t st) {
    if (nseg < 0 || nseg > 31) return fail(P2P_ERR_ARG, "too many halo segments");
    if (nseg == 0) return 0;
    int* hb = reinterpret_cast<int*>(c->h_halo + 32);
    for (int s = 0; s <= nseg; s++) hb[s] = bound[s];
    CU(cudaMemcpyAsync(c->d_halo + 32, hb, (size_t)(nseg + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    p2p::halo_segment_totals_kernel<<<1, 32, 0, st>>>(c->halo_off.p, reinterpret_cast<const int*>(c->d_halo + 32), nseg, c->d_halo);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_halo, c->d_halo, (size_t)nseg * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int s = 0; s < nseg; s++) out[s] = c->h_halo[s];
    return 0;
}
}  // namespace

int p2p_halo_plan_need(p2p_ctx* c, const void* d_topo_all, int npeer, int me, const int* peer_nleaf, int nleaf_max, int nnode_max,
                       void* d_marks, int64_t* need_total) {
    USE(c);
    if (npeer < 1 || npeer > p2p::kHaloPeers || me < 0 || me >= npeer || !peer_nleaf || !need_total || (npeer > 1 && (!d_topo_all || !d_marks)))
        return fail(P2P_ERR_ARG, "bad halo plan arguments");
    cudaStream_t st = c->stream;
    long long off_tb, off_son, off_leaf, stride;
    p2p_topo_layout(nleaf_max, nnode_max, &off_tb, &off_son, &off_leaf, &stride);
    p2p::PeerMap M;
    memset(&M, 0, sizeof M);
    int G = 0, bound[p2p::kHaloPeers + 1];
    for (int p = 0; p < npeer; p++) {
        need_total[p] = 0;
        if (p == me) continue;
        M.first[M.nslot] = G;
        M.leaf_off[M.nslot] = ((long long)p * stride + off_leaf) / 8;
        bound[M.nslot] = G;
        M.nslot++;
        G += peer_nleaf[p];
    }
    M.first[M.nslot] = G;
    bound[M.nslot] = G;
    if (G != c->nghostleaf) return fail(P2P_ERR_STATE, "the peer walk announced %d ghost leaves, the peers hold %d", c->nghostleaf, G);
    c->nghost = 0;
    if (G == 0) return 0;
    CU(c->halo_cnt.reserve((size_t)G + 1, st));
    CU(cudaMemsetAsync(d_marks, 0, (size_t)G, st));
    if (c->ntask) p2p::halo_mark_kernel<<<(unsigned)((c->ntask + 255) / 256), 256, 0, st>>>(c->ts.p, c->ntask, c->nleaf, reinterpret_cast<unsigned char*>(d_marks));
    p2p::halo_need_counts_kernel<<<(G + 255) / 256, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(d_marks), G, M,
                                                                 reinterpret_cast<const int2*>(d_topo_all), c->halo_cnt.p);
    CU(cudaGetLastError());
    int r = halo_scan(c, G, st);
    if (r) return r;
    CU(c->leaf.reserve((size_t)c->nleaf + G + 1, st, (size_t)c->nleaf));
    p2p::halo_ghost_table_kernel<<<(G + 255) / 256, 256, 0, st>>>(c->halo_off.p, c->halo_cnt.p, G, (int)c->npart, kStage, c->leaf.p + c->nleaf,
                                                                 c->d_bad + 1);
    CU(cudaGetLastError());
    long long tot[p2p::kHaloPeers];
    if ((r = halo_totals(c, bound, M.nslot, tot, st))) return r;
    long long sum = 0;
    for (int p = 0, s = 0; p < npeer; p++) {
        if (p == me) continue;
        need_total[p] = tot[s];
        sum += tot[s++];
    }
    if (c->npart + sum > 0x7fffffffLL) return fail(P2P_ERR_ARG, "local + ghost particles exceed 2^31");
    c->nghost = sum;
    c->csr_valid = false;
    c->bounds_n = std::min(c->bounds_n, c->nleaf);
    CU(c->part.reserve((size_t)(c->npart + sum) + 1, st, (size_t)c->npart));
    return 0;
}

int p2p_halo_plan_give(p2p_ctx* c, const void* d_asked, int nreq, int64_t* give_total) {
    USE(c);
    if (nreq < 0 || nreq > 31 || (nreq && (!d_asked || !give_total))) return fail(P2P_ERR_ARG, "bad halo plan arguments");
    cudaStream_t st = c->stream;
    const long long n = (long long)nreq * c->nleaf;
    for (int q = 0; q < nreq; q++) give_total[q] = 0;
    if (n == 0) return 0;
    CU(c->halo_cnt.reserve((size_t)n + 1, st));
    p2p::halo_give_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned char*>(d_asked), n, c->nleaf, c->leaf.p,
                                                                              c->halo_cnt.p);
    CU(cudaGetLastError());
    int r = halo_scan(c, n, st);
    if (r) return r;
    int bound[32];
    for (int q = 0; q <= nreq; q++) bound[q] = q * c->nleaf;
    long long tot[32];
    if ((r = halo_totals(c, bound, nreq, tot, st))) return r;
    for (int q = 0; q < nreq; q++) give_total[q] = tot[q];
    return 0;
}

int p2p_halo_gather(p2p_ctx* c, const void* d_asked, int nreq, void* d_send) {
    USE(c);
    const long long n = (long long)nreq * c->nleaf;
    if (n == 0) return 0;
    if (!d_asked || !d_send) return fail(P2P_ERR_ARG, "null device pointer");
    p2p::halo_gather_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const unsigned char*>(d_asked), n, c->nleaf, c->leaf.p,
                                                                                      c->halo_off.p, c->part.p, reinterpret_cast<int4*>(d_send));
    CU(cudaGetLastError());
    return 0;
}

int p2p_halo_set_particles(p2p_ctx* c, const void* d_recv, int64_t nbody) {
    USE(c);
    if (nbody != c->nghost) return fail(P2P_ERR_STATE, "%lld ghost particles arrived, the plan expects %lld", (long long)nbody, c->nghost);
    if (nbody && !d_recv) return fail(P2P_ERR_ARG, "null device pointer");
    if (nbody) CU(cudaMemcpyAsync(c->part.p + c->npart, d_recv, (size_t)nbody * sizeof(int4), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

void* p2p_device_particles(p2p_ctx* c) { return c ? c->part.p : nullptr; }
void* p2p_device_acc(p2p_ctx* c) { return c ? c->acc.p : nullptr; }

}  // extern "C"
