// p2p_ctx.h -- internal: the context behind include/p2p_b200.h, shared by the translation units of libp2p_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <algorithm>

#include "../../include/p2p_b200.h"

int p2p_fail(int code, const char* fmt, ...);

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return p2p_fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? P2P_ERR_NODEVICE : P2P_ERR_CUDA, \
                        "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);       \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n, cudaStream_t st, size_t keep = 0) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p) {
            e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct p2p_dtree;                       // device-resident kd-tree + walk frontier (device_tree.cu)
void p2p_dtree_release(p2p_dtree* t);

struct p2p_ctx {
    int device = 0;
    int num_sm = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    double mass = 1.0, eps = 0.0, rs = 0.0;
    double origin[3] = {0, 0, 0}, extent = 0.0;   // fixed-point frame; extent 0: derive from the particles
    bool box_set = false;
    int variant = P2P_KERNEL_AUTO;
    int tune_tt = 0, tune_nsrc = 0, tune_minb = 0;
    double far_u = -1.0;                          // near / far threshold in u = r / 2 r_s (< 0: P2P_U_FAR; 0: no far class)
    long long npart = 0, nghost = 0, ntask = 0, npairs = -1;
    // the SECOND list set (tt2 ... below): p2p_swap_lists exchanges the two, so that e.g. the remote (halo) list of a
    // multi-rank step is walked and packed while the force kernel still consumes the local list
    long long ntask_b = 0, npairs_b = -1;
    bool csr_valid_b = false;
    int rows_per_warp = 0;                  // force kernel: 0 = persistent warps, k = a warp retires after k x 2^17 cycles (lets kernels
                                            // of higher-priority streams -- NCCL, the halo walk -- in between)
    int row_lo = 0, row_hi = 0;             // row window of the next packing (row_hi <= row_lo: the whole leaf table), set per chunk by p2p_forces_local
    int nleaf = 0, nghostleaf = 0, max_target_leaf = 0;
    bool csr_valid = false;
    DevBuf<int4> part;
    DevBuf<float4> acc;
    DevBuf<int2> leaf;
    DevBuf<int> tt, ts, col, itmp;
    DevBuf<long long> row_ptr;
    DevBuf<unsigned int> cnt;
    DevBuf<unsigned long long> cursor, tile, row_work, row_work2;
    DevBuf<int> order, order2;
    DevBuf<int> row_mid, row_mid2;          // near columns per row (the rest of the row is far), see csr_pack.cuh
    DevBuf<int4> lbounds;                   // LeafBounds of the leaves [0, bounds_n) (local, then ghost leaves)
    int bounds_n = 0;
    DevBuf<unsigned int> whist, whist2;     // [0,64) histogram of log2(row work), [64,128) bucket cursors
    DevBuf<unsigned char> stage;
    unsigned int* d_counter = nullptr;      // [0] row scheduler, [1] unsorted rows
    unsigned long long* d_npairs = nullptr;      // pairs of the current CSR
    unsigned long long* d_npairs_acc = nullptr;  // pairs accumulated into acc since it was last zeroed
    long long acc_tasks = 0;
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    // second set of list buffers + copy stream for the chunk-pipelined step (p2p_step_host_chunked)
    DevBuf<int> tt2, ts2, col2;
    DevBuf<long long> row_ptr2;
    DevBuf<unsigned int> cnt2;
    DevBuf<unsigned long long> cursor2, tile2;
    unsigned int* d_counter2 = nullptr;
    unsigned int* d_bad = nullptr;          // tasks with ids out of range, summed over every packed list
    unsigned long long* d_npairs2 = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_ready = nullptr, ev_bounds = nullptr;
    unsigned int* h_flags = nullptr;        // pinned copy of d_counter after build_csr
    bool flags_pending = false;
    DevBuf<double> acc64;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    float ms_compute = 0.f, ms_csr = 0.f;
    int last_blocks_per_sm = 0;
    bool timed_compute = false, timed_csr = false;
    cudaEvent_t ev0_b = nullptr, ev1_b = nullptr, ev2_b = nullptr, ev3_b = nullptr;     // ... of the second list set
    float ms_compute_b = 0.f, ms_csr_b = 0.f;
    bool timed_compute_b = false, timed_csr_b = false;
    // halo planning scratch (multi-rank device path, device_tree.cu)
    DevBuf<unsigned int> halo_cnt;
    DevBuf<long long> halo_off;
    DevBuf<unsigned long long> halo_tile, halo_cursor;
    long long* h_halo = nullptr;            // pinned: [0, 32) per-peer totals (out), [32, 64) segment bounds as ints (in)
    long long* d_halo = nullptr;            // the same on the device
    p2p_dtree* dtree = nullptr;
    unsigned int* d_occ = nullptr;          // [64] histogram of the local leaves' occupancies (row schedule band size)
};

// Packed topology block of one rank (p2p_tree_export_packed): [boxes (nleaf_max + nnode_max) x 48 B | tight leaf bounds
// nleaf_max x 48 B | sons nnode_max x 8 B | leaves {first particle, count} nleaf_max x 8 B]; section offsets are fixed by
// the LARGEST tree of the job so that every rank's block has the same stride and the P blocks of one
// all_gather_into_tensor can be walked in place.
inline void p2p_topo_layout(int nlmax, int nnmax, long long* off_tb, long long* off_son, long long* off_leaf, long long* stride) {
    const long long box = ((long long)nlmax + nnmax) * 48;
    *off_tb = box;
    *off_son = box + (long long)nlmax * 48;
    *off_leaf = *off_son + (long long)nnmax * 8;
    const long long end = *off_leaf + (long long)nlmax * 8;
    *stride = (end + 47) / 48 * 48;
}

int p2p_use(p2p_ctx* c);
int p2p_update_occupancy(p2p_ctx* c);     // after the local leaf table changed
#define USE(c)                \
    do {                      \
        int r__ = p2p_use(c); \
        if (r__) return r__;  \
    } while (0)
