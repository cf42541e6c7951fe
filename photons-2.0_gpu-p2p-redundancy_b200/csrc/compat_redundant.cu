// compat_redundant.cu -- the reference's Redundant GPU C-ABI (include/photoNs_CUDA_redundant.h).
//
// The Redundant layout gives every task private fp64 copies of its target and source particles
// (2_Redundant/src/fmm.c:812-838, 2_Redundant/src/remotes.c:55-98).  There is no reuse between
// tasks, so the kernel is a streaming one: one warp per task, lane = target particle, the task's
// sources broadcast from shared memory; coordinates are made relative to the task's first target
// in fp64 before the FP32 arithmetic (exact for the short separations of a leaf pair).  It is
// HBM-bound by construction ((nT+nS)*24 B in, nT*24 B out per task) and exists
// for API parity and for the Indexing-vs-Redundant-vs-CSR layout comparison of BASELINE config 2.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda_runtime.h>

#include "../../include/photoNs_CUDA_redundant.h"
#include "p2p_gcoef.h"

// link-compatibility with the globals the reference header declares extern (cuh:10-21)
double* d_pos_data = nullptr;
double* d_acc_data = nullptr;
int* d_pos_index = nullptr;
double* d_self_pos_data = nullptr;
double* d_self_acc_data = nullptr;
int* d_self_pos_index = nullptr;
int max_res_pos_size_ = 0, max_acc_size = 0, max_posIndexSize = 0;

namespace {

constexpr int kMaxLeaf = 32;

struct TaskDesc { long long tpos, spos, res; int nt, ns; };   // offsets in doubles

struct Phys { float eps2, nlog2e_k2, c[9], out_scale; double scale; int trunc; };

__device__ __forceinline__ float rsqrt_a(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(128) private_tasks_kernel(const double* __restrict__ pos, const TaskDesc* __restrict__ desc,
                                                            int ntask, double* __restrict__ out, Phys P) {
    __shared__ float4 src[4][kMaxLeaf];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int task = blockIdx.x * 4 + w; task < ntask; task += gridDim.x * 4) {
        const TaskDesc d = desc[task];
        if (d.nt <= 0 || d.ns <= 0) continue;
        const double ox = pos[d.tpos], oy = pos[d.tpos + 1], oz = pos[d.tpos + 2];
        __syncwarp();
        if (lane < d.ns) {
            const double* p = pos + d.spos + 3 * lane;
            src[w][lane] = make_float4((float)((p[0] - ox) * P.scale), (float)((p[1] - oy) * P.scale),
                                       (float)((p[2] - oz) * P.scale), 0.f);
        }
        float tx = 0.f, ty = 0.f, tz = 0.f;
        if (lane < d.nt) {
            const double* p = pos + d.tpos + 3 * lane;
            tx = (float)((p[0] - ox) * P.scale); ty = (float)((p[1] - oy) * P.scale); tz = (float)((p[2] - oz) * P.scale);
        }
        __syncwarp();
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int j = 0; j < d.ns; j++) {
            const float4 s = src[w][j];
            const float dx = s.x - tx, dy = s.y - ty, dz = s.z - tz;
            float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            r2 = fmaxf(r2, P.eps2);
            const float rinv = rsqrt_a(r2), rinv2 = rinv * rinv;
            float f;
            if (P.trunc) {
                const float e = ex2_a(r2 * P.nlog2e_k2), v = r2 * rinv;
                float R = P.c[8];
#pragma unroll
                for (int k = 7; k >= 0; k--) R = fmaf(R, v, P.c[k]);
                f = (rinv2 * e) * fmaf(v, R, rinv);
            } else {
                f = rinv2 * rinv;
            }
            ax = fmaf(dx, f, ax); ay = fmaf(dy, f, ay); az = fmaf(dz, f, az);
        }
        if (lane < d.nt) {
            double* o = out + d.res + 3 * lane;
            o[0] = (double)(ax * P.out_scale); o[1] = (double)(ay * P.out_scale); o[2] = (double)(az * P.out_scale);
        }
    }
}

double g_rs = 0.0;
bool g_env_checked = false, g_init = false, g_have_device = false;
int g_max_parts = 0, g_maxtask = 0;

struct Batch {                 // one uploaded batch of private-copy tasks
    double* d_pos = nullptr; size_t cap_pos = 0;
    TaskDesc* d_desc = nullptr; size_t cap_desc = 0;
    double* d_out = nullptr; size_t cap_out = 0;
    size_t n_out = 0; int ntask = 0; bool ok = false;
} g_remote, g_self;

void check_env() {
    if (g_env_checked) return;
    g_env_checked = true;
    const char* e = getenv("P2P_B200_RS");
    if (e && g_rs == 0.0) g_rs = atof(e);
}

template <typename T>
bool ensure(T** p, size_t* cap, size_t n) {
    if (n <= *cap) return true;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t want = n + n / 2 + 16;
    if (cudaMalloc(p, want * sizeof(T)) != cudaSuccess) { printf("Error allocating %zu bytes on the GPU : %s\n", want * sizeof(T), cudaGetErrorString(cudaGetLastError())); return false; }
    *cap = want;
    return true;
}

bool upload(Batch& b, const double* h_pos, size_t npos, const std::vector<TaskDesc>& desc, size_t nout) {
    b.ok = false;
    if (!ensure(&b.d_pos, &b.cap_pos, npos + 1) || !ensure(&b.d_desc, &b.cap_desc, desc.size() + 1) || !ensure(&b.d_out, &b.cap_out, nout + 1)) return false;
    if (npos && cudaMemcpy(b.d_pos, h_pos, npos * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (!desc.empty() && cudaMemcpy(b.d_desc, desc.data(), desc.size() * sizeof(TaskDesc), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (cudaMemset(b.d_out, 0, (nout + 1) * sizeof(double)) != cudaSuccess) return false;
    b.n_out = nout; b.ntask = (int)desc.size(); b.ok = true;
    return true;
}

int launch(Batch& b, double eps, double mass) {
    if (!b.ok) return -1;
    check_env();
    Phys P;
    memset(&P, 0, sizeof P);
    P.trunc = g_rs > 0.0;
    P.scale = P.trunc ? exp2(round(log2(1.0 / (2.0 * g_rs)))) : 1.0;
    P.eps2 = (float)((eps * P.scale) * (eps * P.scale));
    if (P.trunc) {
        const double kappa = 1.0 / (2.0 * g_rs * P.scale);
        P.nlog2e_k2 = (float)(-1.4426950408889634 * kappa * kappa);
        for (int j = 0; j < 9; j++) P.c[j] = (float)(P2P_GCOEF_10[j + 2] * pow(kappa, j + 2));
    }
    P.out_scale = (float)(P.scale * P.scale * mass);
    if (b.ntask > 0) {
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
        int grid = (b.ntask + 3) / 4;
        if (grid > sm * 16) grid = sm * 16;
        private_tasks_kernel<<<grid, 128>>>(b.d_pos, b.d_desc, b.ntask, b.d_out, P);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error kernel : %s\n", cudaGetErrorString(e)); return -1; }
    return 0;
}

}  // namespace

extern "C" {

void p2pSetSplitRadius(double rs) { g_rs = rs; g_env_checked = true; }

void initGPU(int verbosity_gpu) {
    if (g_init) return;
    g_init = true;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { printf("No CUDA-capable device found!\n"); cudaGetLastError(); return; }
    cudaSetDevice(0);
    g_have_device = true;
    if (verbosity_gpu) printf(">> \tp2p_b200 (redundant ABI) on device 0\n");
}

void getGPUMemoryState(int verbosity_gpu) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { if (verbosity_gpu) printf("cudaMemGetInfo failed\n"); return; }
    if (verbosity_gpu) printf(">> \tFree memory: %zu bytes\n", fr);
}

int allocMemGPU(int PROC_SIZE, int maxPartsInLeaf, int MAXTASK, int verbosity_gpu) {
    (void)PROC_SIZE;
    initGPU(verbosity_gpu);
    if (!g_have_device) return -1;
    cudaSetDevice(0);
    g_max_parts = maxPartsInLeaf; g_maxtask = MAXTASK;
    return 0;
}

int copyMemGPU(double** h_pos_data, int** h_pos_index, int PROC_SIZE, int PROC_RANK, int maxPartsInLeaf, int numtasks,
               int posCounter, int verbosity_gpu) {
    (void)PROC_SIZE; (void)verbosity_gpu;
    g_remote.ok = false;
    if (!g_have_device) { printf("Error copy for d_pos_data : no device\n"); return -1; }
    if (!h_pos_data || !h_pos_index || numtasks < 0 || posCounter < 0 || maxPartsInLeaf <= 0 || maxPartsInLeaf > kMaxLeaf) { printf("Error copy for d_pos_data : bad arguments\n"); return -1; }
    cudaSetDevice(0);
    const int* idx = h_pos_index[PROC_RANK];
    std::vector<TaskDesc> desc((size_t)numtasks);
    size_t nout = 0;
    for (int n = 0; n < numtasks; n++) {
        const int start = idx[5 * n], nt = idx[5 * n + 2], ns = idx[5 * n + 3], res = idx[5 * n + 4];
        if (nt < 0 || ns < 0) { printf("ERROR : task %d has negative counts\n", n); return -1; }          // kernel code -1, cu:243-248
        if (nt > kMaxLeaf || ns > kMaxLeaf || start < 0 || res < 0 || (long long)start + 3LL * (nt + ns) > posCounter) {
            printf("ERROR : task %d offsets outside the uploaded data\n", n);                                 // kernel code -3, cu:253-262
            return -3;
        }
        desc[(size_t)n] = TaskDesc{start, (long long)start + 3LL * nt, res, nt, ns};
        if ((size_t)res + 3 * (size_t)nt > nout) nout = (size_t)res + 3 * (size_t)nt;
    }
    if (!upload(g_remote, h_pos_data[PROC_RANK], (size_t)posCounter, desc, nout)) { printf("Error copy for d_pos_data\n"); return -1; }
    return 0;
}

int LaunchKernelP2PDualNaive(int PROC_SIZE, int PROC_RANK, int nTasks, double SoftenScale, double MASSPART, int verbosity_gpu) {
    (void)PROC_SIZE; (void)PROC_RANK; (void)verbosity_gpu;
    if (!g_remote.ok || nTasks != g_remote.ntask) { printf("error kernel ComputeP2PDualNaive : no matching upload\n"); return -1; }
    return launch(g_remote, SoftenScale, MASSPART);
}

void readResultsGPU(double** h_acc_data, int PROC_RANK, int PROC_SIZE, int maxPartsInLeaf, int MAXTASK, int partCounter,
                    int verbosity_gpu) {
    (void)PROC_SIZE; (void)maxPartsInLeaf; (void)MAXTASK; (void)verbosity_gpu;
    if (!g_remote.ok || !h_acc_data || partCounter < 0) { printf("Error copy for reading results (h_acc_data)\n"); return; }
    size_t n = (size_t)partCounter < g_remote.n_out ? (size_t)partCounter : g_remote.n_out;
    if (n && cudaMemcpy(h_acc_data[PROC_RANK], g_remote.d_out, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        printf("Error copy for reading results (h_acc_data)\n");
    if ((size_t)partCounter > n) memset(h_acc_data[PROC_RANK] + n, 0, sizeof(double) * ((size_t)partCounter - n));
}

int allocAndCopySelfInteractionsGPU(double* part_data, int* part_idx, int partDataChunk, int partIndexChunk,
                                    int resultDataChunk, int nTasks) {
    g_self.ok = false;
    initGPU(0);
    if (!g_have_device) return -1;
    if (!part_data || !part_idx || partIndexChunk < 2 || nTasks < 0 || partDataChunk <= 0 || resultDataChunk <= 0) return -1;
    cudaSetDevice(0);
    std::vector<TaskDesc> desc((size_t)nTasks);
    for (int n = 0; n < nTasks; n++) {
        const int nt = part_idx[n * partIndexChunk], ns = part_idx[n * partIndexChunk + 1];
        if (nt < 0 || ns < 0 || nt > kMaxLeaf || ns > kMaxLeaf || 3 * (nt + ns) > partDataChunk || 3 * nt > resultDataChunk) {
            printf("ERROR : self task %d does not fit its chunks (nT %d nS %d)\n", n, nt, ns);
            return -1;
        }
        const long long base = (long long)n * partDataChunk;
        desc[(size_t)n] = TaskDesc{base, base + 3LL * nt, (long long)n * resultDataChunk, nt, ns};
    }
    if (!upload(g_self, part_data, (size_t)nTasks * (size_t)partDataChunk, desc, (size_t)nTasks * (size_t)resultDataChunk)) return -1;
    return 0;
}

void LaunchKernelP2PSelfInteractions(int nTasks, int partDataChunk, int partIndexChunk, int resultDataChunk, double SoftenScale,
                                     double MASSPART) {
    (void)partDataChunk; (void)partIndexChunk; (void)resultDataChunk;
    if (!g_self.ok || nTasks != g_self.ntask) { printf("ERROR : kernel ComputeP2PSelfInteractions : no matching upload\n"); return; }
    launch(g_self, SoftenScale, MASSPART);
}

void readResultsGPUSelfInteractions(double* h_acc_data, int accDataChunk, int nTasks) {
    if (!g_self.ok || !h_acc_data) { printf("memcpy error for d_self_acc_data\n"); return; }
    size_t n = (size_t)accDataChunk * (size_t)nTasks;
    if (n > g_self.n_out) n = g_self.n_out;
    if (n && cudaMemcpy(h_acc_data, g_self.d_out, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        printf("memcpy error for d_self_acc_data\n");
}

}  // extern "C"
