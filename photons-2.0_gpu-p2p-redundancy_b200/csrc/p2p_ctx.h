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
    int tune_tt = 0, tune_nsrc = 0, tune_minb = 0, tune_poly = 0;
    long long npart = 0, nghost = 0, ntask = 0, npairs = -1;
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
    DevBuf<unsigned int> whist, whist2;     // [0,64) histogram of log2(row work), [64,128) bucket cursors
    DevBuf<unsigned char> stage;
    unsigned int* d_counter = nullptr;      // [0] row scheduler, [1] unsorted rows
    unsigned long long* d_npairs = nullptr;      // [0] pairs of the current CSR, [1] pairs accumulated into acc
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
    cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_ready = nullptr;
    unsigned int* h_flags = nullptr;        // pinned copy of d_counter after build_csr
    bool flags_pending = false;
    DevBuf<double> acc64;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    float ms_compute = 0.f, ms_csr = 0.f;
    int last_blocks_per_sm = 0;
    bool timed_compute = false, timed_csr = false;
    p2p_dtree* dtree = nullptr;
    unsigned int* d_occ = nullptr;          // [64] histogram of the local leaves' occupancies (row schedule band size)
};

int p2p_use(p2p_ctx* c);
int p2p_update_occupancy(p2p_ctx* c);     // after the local leaf table changed
#define USE(c)                \
    do {                      \
        int r__ = p2p_use(c); \
        if (r__) return r__;  \
    } while (0)
