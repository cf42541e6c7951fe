// Micro-benchmark of the sm_100a CUDA-core pipes that bound the P2P pair kernel.
// Measures warp-instruction issue throughput (lane-ops per clock per SM) for
// FFMA (3-register), packed FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2), FMNMX,
// MUFU.RSQ, MUFU.EX2 and two instruction mixes that resemble the pair kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 65536
#define NCHAIN 8

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, float seed, long long* cycles) {
    float a[NCHAIN], b = seed, c = seed * 0.5f;
    float2 a2[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) { a[i] = seed + i + threadIdx.x; a2[i] = make_float2(a[i], a[i] + 1.f); }
    float2 b2 = make_float2(b, b + 0.25f), c2 = make_float2(c, c + 0.125f);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                       // FFMA
            if (MODE == 1) a2[i] = __ffma2_rn(a2[i], b2, c2);             // FFMA2
            if (MODE == 2) a2[i] = __fmul2_rn(a2[i], b2);                 // FMUL2
            if (MODE == 3) a2[i] = __fadd2_rn(a2[i], b2);                 // FADD2
            if (MODE == 4) a[i] = fmaxf(a[i], a[(i + 1) % NCHAIN]) ;      // FMNMX (data dependent)
            if (MODE == 5) a[i] = rsqrtf(a[i]);                           // MUFU.RSQ
            if (MODE == 6) a[i] = exp2f(a[i]);                            // MUFU.EX2
            if (MODE == 7) { a[i] = fmaf(a[i], b, c); a[i] = a[i] * b; a[i] = a[i] + c; }  // FFMA+FMUL+FADD
            if (MODE == 8) {  // scalar mix: 6 FFMA + 1 FMNMX + 1 MUFU per 8
                a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c);
                a[i] = fmaxf(a[i], c); a[i] = rsqrtf(a[i]);
                a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c);
            }
            if (MODE == 9) {  // packed mix: 6 FFMA2 + 2 FMNMX + 2 MUFU per 2x8 lane-ops
                a2[i] = __ffma2_rn(a2[i], b2, c2); a2[i] = __ffma2_rn(a2[i], b2, c2); a2[i] = __ffma2_rn(a2[i], b2, c2);
                a2[i].x = fmaxf(a2[i].x, c); a2[i].y = fmaxf(a2[i].y, c);
                a2[i].x = rsqrtf(a2[i].x); a2[i].y = rsqrtf(a2[i].y);
                a2[i] = __ffma2_rn(a2[i], b2, c2); a2[i] = __ffma2_rn(a2[i], b2, c2); a2[i] = __ffma2_rn(a2[i], b2, c2);
            }
            if (MODE == 12) a2[i] = __ffma2_rn(a2[i], a2[(i + 1) % NCHAIN], a2[(i + 3) % NCHAIN]);   // FFMA2, 3 distinct register pairs
            if (MODE == 13) a[i] = fmaf(a[i], a[(i + 1) % NCHAIN], a[(i + 3) % NCHAIN]);            // FFMA, 3 distinct registers
            if (MODE == 14) {   // pair-kernel-like packed mix: 11 FFMA2-class (distinct operands) + 1 FMNMX pair + 2 MUFU pairs per 2x... 
                float2 t = __ffma2_rn(a2[i], a2[(i + 1) % NCHAIN], a2[(i + 3) % NCHAIN]);
                t = __ffma2_rn(t, a2[(i + 2) % NCHAIN], b2); t = __ffma2_rn(t, a2[(i + 5) % NCHAIN], c2);
                t.x = fmaxf(t.x, c); t.y = fmaxf(t.y, c);
                float2 r = make_float2(rsqrtf(t.x), rsqrtf(t.y));
                float2 e = make_float2(exp2f(-t.x), exp2f(-t.y));
                float2 w = __fmul2_rn(t, r);
                float2 E = __ffma2_rn(w, t, b2), O = __ffma2_rn(w, t, c2);
                E = __ffma2_rn(E, t, c2); O = __ffma2_rn(O, t, b2); E = __ffma2_rn(E, t, b2); O = __ffma2_rn(O, t, c2);
                E = __ffma2_rn(w, O, E); E = __ffma2_rn(w, E, r);
                r = __fmul2_rn(r, r); e = __fmul2_rn(r, e); e = __fmul2_rn(e, E);
                a2[i] = __ffma2_rn(a2[(i + 1) % NCHAIN], e, a2[i]);
            }
            if (MODE >= 20 && MODE <= 24) {   // 10 reuse-friendly FFMA2 + (MODE-20) MUFU on a separate chain: pipe interference
#pragma unroll
                for (int k = 0; k < 10; k++) a2[i] = __ffma2_rn(a2[i], b2, c2);
                if (MODE >= 21) a[i] = rsqrtf(a[i]);
                if (MODE >= 22) a[i] = exp2f(a[i]);
                if (MODE >= 23) a[i] = rsqrtf(a[i]);
                if (MODE >= 24) a[i] = exp2f(a[i]);
            }
            if (MODE >= 25 && MODE <= 27) {   // 10 FFMA2 + (MODE-25)*2 FMNMX
#pragma unroll
                for (int k = 0; k < 10; k++) a2[i] = __ffma2_rn(a2[i], b2, c2);
                if (MODE >= 26) { a[i] = fmaxf(a[i], a[(i + 1) % NCHAIN]); a[i] = fminf(a[i], a[(i + 2) % NCHAIN]); }
                if (MODE >= 27) { a[i] = fmaxf(a[i], a[(i + 3) % NCHAIN]); a[i] = fminf(a[i], a[(i + 5) % NCHAIN]); }
            }
            if (MODE == 10) a[i] = a[i] * b;                              // FMUL
            if (MODE == 11) a[i] = a[i] + b;                              // FADD
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) s += a[i] + a2[i].x + a2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter_per_thread, int nsm, float* d_out, long long* d_cyc) {
    int blocks = nsm * 4, threads = 256;  // 32 warps/SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<blocks, threads>>>(d_out, 1.0001f, d_cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<MODE><<<blocks, threads>>>(d_out, 1.0001f, d_cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[4096]; cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += cyc[i]; avg /= blocks;
    double ops_per_sm = lane_ops_per_iter_per_thread * ITERS * NCHAIN * threads * 4;  // 4 blocks/SM
    printf("%-28s ms=%.4f  cyc/block=%.0f  lane-ops/clk/SM=%.1f  (clk est %.0f MHz)\n", name, ms, avg,
           ops_per_sm / avg, avg / (ms * 1e3));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("device %s  SMs=%d  clock=%d kHz\n", p.name, nsm, p.clockRate);
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, sizeof(float) * nsm * 4 * 256); cudaMalloc(&d_cyc, sizeof(long long) * 4096);
    for (int w = 0; w < 200; w++) kern<0><<<nsm * 4, 256>>>(d_out, 1.0001f, d_cyc);
    cudaDeviceSynchronize();
    run<0>("FFMA", 1, nsm, d_out, d_cyc);
    run<10>("FMUL", 1, nsm, d_out, d_cyc);
    run<11>("FADD", 1, nsm, d_out, d_cyc);
    run<1>("FFMA2 (2 lane-ops)", 2, nsm, d_out, d_cyc);
    run<2>("FMUL2", 2, nsm, d_out, d_cyc);
    run<3>("FADD2", 2, nsm, d_out, d_cyc);
    run<4>("FMNMX", 1, nsm, d_out, d_cyc);
    run<5>("MUFU.RSQ", 1, nsm, d_out, d_cyc);
    run<6>("MUFU.EX2", 1, nsm, d_out, d_cyc);
    run<7>("FFMA+FMUL+FADD", 3, nsm, d_out, d_cyc);
    run<8>("mix scalar 6F+MNMX+RSQ", 8, nsm, d_out, d_cyc);
    run<9>("mix packed 6F2+2MNMX+2RSQ", 16, nsm, d_out, d_cyc);
    run<20>("10 FFMA2", 20, nsm, d_out, d_cyc);
    run<21>("10 FFMA2 + 1 MUFU", 20, nsm, d_out, d_cyc);
    run<22>("10 FFMA2 + 2 MUFU", 20, nsm, d_out, d_cyc);
    run<23>("10 FFMA2 + 3 MUFU", 20, nsm, d_out, d_cyc);
    run<24>("10 FFMA2 + 4 MUFU", 20, nsm, d_out, d_cyc);
    run<25>("10 FFMA2 + 0 FMNMX", 20, nsm, d_out, d_cyc);
    run<26>("10 FFMA2 + 2 FMNMX", 20, nsm, d_out, d_cyc);
    run<27>("10 FFMA2 + 4 FMNMX", 20, nsm, d_out, d_cyc);
    run<12>("FFMA2 3 distinct reg pairs", 2, nsm, d_out, d_cyc);
    run<13>("FFMA 3 distinct regs", 1, nsm, d_out, d_cyc);
    run<14>("pair-like packed (17 F2 + 2 MNMX + 4 MUFU)", 34, nsm, d_out, d_cyc);
    return 0;
}
