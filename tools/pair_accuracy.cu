// pair_accuracy.cu -- where does the FP32 pair arithmetic lose accuracy?  (development aid, standalone)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o tools/pair_accuracy tools/pair_accuracy.cu
// For random separations in bins of u = r / 2 r_s, evaluates the near-field force factor of csrc/p2p_kernel.cuh
// (rinv^3 2^(-w) Q) in FP32 with each approximate unit (MUFU.RSQ, MUFU.EX2) replaced in turn by the correctly rounded
// value, and prints the rms / max relative error against fp64 erfc / exp.  Also the raw error of the three MUFU ops.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../photons-2.0_gpu-p2p-redundancy_b200/csrc/p2p_gcoef.h"

struct Coef { float c[9]; };

__device__ double g_exact(double u) { return erfc(u) + 1.1283791670955126 * u * exp(-u * u); }

// mode bit 0: exact rsqrt, bit 1: exact exp2, bit 2: Newton step on the approximate rsqrt
__device__ float pair_f(float r2, const Coef& K, int mode) {
    float rinv = (mode & 1) ? (float)(1.0 / sqrt((double)r2)) : rsqrtf(r2);
    const float e = (mode & 2) ? (float)exp2(-(double)r2) : exp2f(-r2);
    const float v = r2 * rinv;
    if (mode & 4) { const float d = fmaf(v, rinv, -1.f); rinv = fmaf(rinv * d, -0.5f, rinv); }
    float E = K.c[8], O = K.c[7];
    E = fmaf(E, r2, K.c[6]); O = fmaf(O, r2, K.c[5]);
    E = fmaf(E, r2, K.c[4]); O = fmaf(O, r2, K.c[3]);
    E = fmaf(E, r2, K.c[2]); O = fmaf(O, r2, K.c[1]);
    E = fmaf(E, r2, fmaf(rinv, rinv, K.c[0]));
    return (e * rinv) * fmaf(v, O, E);
}

__global__ void run(Coef K, double ulo, double uhi, int n, int mode, double* sum2, double* mx, double* bias) {
    const double c2 = 1.4426950408889634;
    double s2 = 0, m = 0, b = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const double u = ulo + (uhi - ulo) * ((h & 0xffffff) + 0.5) / 16777216.0;
        const float r2 = (float)(u * u * c2);
        const double ue = sqrt((double)r2 / c2);                       // the u the kernel actually sees
        const double want = g_exact(ue) / ((double)r2 * sqrt((double)r2));
        const double rel = (double)pair_f(r2, K, mode) / want - 1.0;
        s2 += rel * rel; b += rel; m = fmax(m, fabs(rel));
    }
    atomicAdd(sum2, s2); atomicAdd(bias, b);
    unsigned long long* a = (unsigned long long*)mx;
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double(assumed) >= m) break; old = atomicCAS(a, assumed, __double_as_longlong(m)); } while (assumed != old);
}

__global__ void mufu(double lo, double hi, int n, int which, double* sum2, double* mx) {
    double s2 = 0, m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = (float)(lo + (hi - lo) * (i + 0.5) / n);
        float y; double w;
        if (which == 0) { y = rsqrtf(x); w = 1.0 / sqrt((double)x); }
        else if (which == 1) { y = exp2f(-x); w = exp2(-(double)x); }
        else { asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); w = 1.0 / (double)x; }
        const double rel = (double)y / w - 1.0;
        s2 += rel * rel; m = fmax(m, fabs(rel));
    }
    atomicAdd(sum2, s2);
    unsigned long long* a = (unsigned long long*)mx;
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double(assumed) >= m) break; old = atomicCAS(a, assumed, __double_as_longlong(m)); } while (assumed != old);
}

int main() {
    Coef K;
    const double sl2e = sqrt(1.4426950408889634);
    for (int j = 0; j < 9; j++) K.c[j] = (float)(P2P_GCOEF_10[j + 2] / pow(sl2e, j + 2));
    double* d;
    cudaMalloc(&d, 3 * sizeof(double));
    const int n = 1 << 24;
    const char* names[3] = {"rsqrt.approx", "ex2.approx(-x)", "rcp.approx"};
    for (int w = 0; w < 3; w++) {
        cudaMemset(d, 0, 3 * sizeof(double));
        mufu<<<592, 256>>>(0.01, 30.0, n, w, d, d + 1);
        double h[3];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("%-16s x in [0.01, 30]: rms rel %.3e  max rel %.3e\n", names[w], sqrt(h[0] / n), h[1]);
    }
    const double bins[][2] = {{0.02, 0.25}, {0.25, 0.5}, {0.5, 1.0}, {1.0, 1.5}, {1.5, 2.0}, {2.0, 2.5}};
    const char* modes[] = {"all approximate", "exact rsqrt", "exact exp2", "both exact", "Newton rsqrt", "-", "Newton + exact exp2"};
    for (auto& b : bins)
        for (int mode : {0, 1, 2, 3, 4, 6}) {
            cudaMemset(d, 0, 3 * sizeof(double));
            run<<<592, 256>>>(K, b[0], b[1], n, mode, d, d + 1, d + 2);
            double h[3];
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("u in [%.2f, %.2f) %-20s rms rel %.3e  max %.3e  mean %+.3e\n", b[0], b[1], modes[mode], sqrt(h[0] / n), h[1], h[2] / n);
        }
    return 0;
}
