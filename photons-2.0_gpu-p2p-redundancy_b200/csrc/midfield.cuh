// midfield.cuh -- the mid-field half of the short-range FMM step on the device:
//   P2M / M2M   1_Indexing/src/operator.c:13-93, 96-160, 165-194 (fmm_prepare, 1_Indexing/src/fmm.c:745-790)
//   M2L         1_Indexing/src/operator.c:255-392 with the erfc-split radial factors :296-305; task lists from
//               walk_task_m2l (1_Indexing/src/fmm.c:562-705) and walk_task_m2l_ext (1_Indexing/src/remotes.c:477-640)
//   L2L / L2P   1_Indexing/src/operator.c:395-494, 498-530, 197-251 (tail of fmm_ext, 1_Indexing/src/fmm.c:1121-1128)
// Cartesian Taylor expansions to third order (the reference compiles QUADRUPOLE + OCTUPOLE: 20 coefficients), in fp64
// like the reference.  A coefficient with multi-index n = (a, b, c) carries 1 / (a! b! c!):
//   M_n  = sum_particles m (-d)^n / n!          d = x - centre
//   M'_n = sum_{k <= n} M_k s^(n-k) / (n-k)!    s = new centre - old centre                      (M2M)
//   L_n += sum_{|n|+|k| <= 3} M_k D_{n+k}(r)    D = derivatives of phi(r) = erfc(r / 2 r_s) / r   (M2L)
//   L'_n += sum_{k >= n} L_k s^(k-n) / (k-n)!   s = new centre - old centre                      (L2L)
//   a_i  = sum_{|k| <= 2} L_{k + e_i} d^k / k!                                                   (L2P)
// written once over multi-index tables (the reference spells every term out); the index order is the reference's:
// 0 | X Y Z | XX XY XZ YY YZ ZZ | XXX XXY XXZ XYY XYZ XZZ YYY YYZ YZZ ZZZ  (1_Indexing/inc/operator.h:24-67).
#pragma once
#include <cuda_runtime.h>

namespace p2p {
namespace mf {

constexpr int NM = 20;

struct Idx { int a, b, c; };
__host__ __device__ constexpr Idx idx_of(int n) {
    constexpr int T[NM][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {0, 2, 0}, {0, 1, 1}, {0, 0, 2},
                              {3, 0, 0}, {2, 1, 0}, {2, 0, 1}, {1, 2, 0}, {1, 1, 1}, {1, 0, 2}, {0, 3, 0}, {0, 2, 1}, {0, 1, 2}, {0, 0, 3}};
    return Idx{T[n][0], T[n][1], T[n][2]};
}
__host__ __device__ constexpr int order_of(int n) { return idx_of(n).a + idx_of(n).b + idx_of(n).c; }
// position of (a, b, c) in the coefficient array, -1 beyond third order
__host__ __device__ constexpr int pos_of(int a, int b, int c) {
    for (int n = 0; n < NM; n++)
        if (idx_of(n).a == a && idx_of(n).b == b && idx_of(n).c == c) return n;
    return -1;
}
__host__ __device__ constexpr double fact(int n) { return n <= 1 ? 1.0 : (n == 2 ? 2.0 : 6.0); }

// pw[n] = x^a y^b z^c / (a! b! c!)
__device__ __forceinline__ void powers(double x, double y, double z, double pw[NM]) {
#pragma unroll
    for (int n = 0; n < NM; n++) {
        const Idx I = idx_of(n);
        double v = 1.0;
        for (int k = 0; k < I.a; k++) v *= x;
        for (int k = 0; k < I.b; k++) v *= y;
        for (int k = 0; k < I.c; k++) v *= z;
        pw[n] = v / (fact(I.a) * fact(I.b) * fact(I.c));
    }
}

// out_n += sum_{k <= n} in_k pw_{n-k}      (M2M with pw of the centre shift)
__device__ __forceinline__ void shift_up(const double in[NM], const double pw[NM], double out[NM]) {
#pragma unroll
    for (int n = 0; n < NM; n++) {
        const Idx N = idx_of(n);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < NM; k++) {
            const Idx K = idx_of(k);
            if (K.a <= N.a && K.b <= N.b && K.c <= N.c) s += in[k] * pw[pos_of(N.a - K.a, N.b - K.b, N.c - K.c)];
        }
        out[n] += s;
    }
}

// out_n += sum_{k >= n} in_k pw_{k-n}      (L2L with pw of the centre shift)
__device__ __forceinline__ void shift_down(const double in[NM], const double pw[NM], double out[NM]) {
#pragma unroll
    for (int n = 0; n < NM; n++) {
        const Idx N = idx_of(n);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < NM; k++) {
            const Idx K = idx_of(k);
            if (K.a >= N.a && K.b >= N.b && K.c >= N.c) s += in[k] * pw[pos_of(K.a - N.a, K.b - N.b, K.c - N.c)];
        }
        out[n] += s;
    }
}

// radial factors f_q with d^q phi / (r dr)^q structure: D_0 = f0, D_i = f1 x_i, D_ij = f2 x_i x_j + f1 delta_ij, ...
// (1_Indexing/src/operator.c:283-305); rs <= 0: plain 1 / r
__device__ __forceinline__ void radial_factors(double r2, double rs, double f[4]) {
    const double dr = sqrt(r2);
    const double ir = 1.0 / dr, ir2 = ir * ir, ir3 = ir2 * ir, ir4 = ir3 * ir, ir5 = ir4 * ir, ir6 = ir5 * ir, ir7 = ir6 * ir;
    if (rs > 0.0) {
        const double irs = 1.0 / rs, irs2 = irs * irs, irs3 = irs2 * irs, irs5 = irs3 * irs2;
        const double u = 0.5 * dr / rs;
        const double fe = exp(-u * u) * 0.56418958354775628695;      // 1 / sqrt(pi)
        const double fc = erfc(u);
        f[0] = ir * fc;
        f[1] = -ir3 * (fc + dr * fe * irs);
        f[2] = 3.0 * ir5 * fc + (3.0 * irs * ir4 + 0.5 * ir2 * irs3) * fe;
        f[3] = -15.0 * ir7 * fc - (15.0 * ir6 * irs + 2.5 * ir4 * irs3 + 0.25 * ir2 * irs5) * fe;
    } else {
        f[0] = ir; f[1] = -ir3; f[2] = 3.0 * ir5; f[3] = -15.0 * ir7;
    }
}

// derivative tensor of the radial function at x (all 20 components)
__device__ __forceinline__ void derivative_tensor(const double x[3], const double f[4], double D[NM]) {
    D[0] = f[0];
#pragma unroll
    for (int n = 1; n < NM; n++) {
        const Idx N = idx_of(n);
        const int ord = N.a + N.b + N.c;
        // the axes of the multi-index, e.g. (2,0,1) -> 0,0,2
        int ax[3] = {0, 0, 0};
        int m = 0;
        for (int k = 0; k < N.a; k++) ax[m++] = 0;
        for (int k = 0; k < N.b; k++) ax[m++] = 1;
        for (int k = 0; k < N.c; k++) ax[m++] = 2;
        if (ord == 1) D[n] = f[1] * x[ax[0]];
        else if (ord == 2) D[n] = f[2] * x[ax[0]] * x[ax[1]] + (ax[0] == ax[1] ? f[1] : 0.0);
        else {
            double t = 0.0;
            if (ax[1] == ax[2]) t += x[ax[0]];
            if (ax[0] == ax[2]) t += x[ax[1]];
            if (ax[0] == ax[1]) t += x[ax[2]];
            D[n] = f[3] * x[ax[0]] * x[ax[1]] * x[ax[2]] + f[2] * t;
        }
    }
}

// ------------------------------------------------------------------------------------------------ kernels
// one thread per leaf: multipole about the leaf's kd-cell centre, particles in tree order (1_Indexing/src/fmm.c:783)
__global__ void p2m_kernel(const int2* __restrict__ leaf, int nleaf, const double* __restrict__ box, const double* __restrict__ px,
                           const double* __restrict__ py, const double* __restrict__ pz, double mass, double* __restrict__ Mall) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaf) return;
    const int2 L = leaf[l];
    const double cx = box[6 * (size_t)l], cy = box[6 * (size_t)l + 1], cz = box[6 * (size_t)l + 2];
    double M[NM];
#pragma unroll
    for (int n = 0; n < NM; n++) M[n] = 0.0;
    for (int p = L.x; p < L.x + L.y; p++) {
        double pw[NM];
        powers(px[p] - cx, py[p] - cy, pz[p] - cz, pw);
#pragma unroll
        for (int n = 0; n < NM; n++) M[n] += (order_of(n) & 1) ? -mass * pw[n] : mass * pw[n];
    }
#pragma unroll
    for (int n = 0; n < NM; n++) Mall[NM * (size_t)l + n] = M[n];
}

// one thread per node of one tree level, deepest level first: M = shifted M of son 0, then of son 1 (walk_m2m)
__global__ void m2m_level_kernel(const int* __restrict__ t_id, int lvl_begin, int lvl_count, const int* __restrict__ son, int nleaf,
                                 const double* __restrict__ box, double* __restrict__ Mall) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_count) return;
    const int id = t_id[lvl_begin + i];
    const size_t u = (size_t)nleaf + id;
    double M[NM];
#pragma unroll
    for (int n = 0; n < NM; n++) M[n] = 0.0;
    for (int s = 0; s < 2; s++) {
        const int c = son[2 * id + s];
        if (c < 0) continue;
        double in[NM], pw[NM];
#pragma unroll
        for (int n = 0; n < NM; n++) in[n] = Mall[NM * (size_t)c + n];
        powers(box[6 * u] - box[6 * (size_t)c], box[6 * u + 1] - box[6 * (size_t)c + 1], box[6 * u + 2] - box[6 * (size_t)c + 2], pw);
        shift_up(in, pw, M);
    }
#pragma unroll
    for (int n = 0; n < NM; n++) Mall[NM * u + n] = M[n];
}

struct M2LParams {
    const int* mt;            // target: local unified id (leaf or node)
    const int* ms;            // source: unified id inside its rank's tree
    const int* mq;            // (rank << 5) | displacement index
    long long ntask;
    const double* box;        // local tree
    const double* sbox;       // source trees (concatenated, as in the walk)
    const double* sM;         // source multipoles (concatenated like sbox, NM per id)
    long long sbase[16];      // first box of rank p in sbox
    long long sbase_M[16];    // first multipole of rank p in sM
    double period, rs;
    double* L;                // [local unified id][NM], accumulated with atomics
};

__constant__ int c_mshift[28][3];

// one thread per M2L task (task_compute_m2l / task_compute_m2l_ext: dx = target centre - source centre)
__global__ void m2l_kernel(M2LParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.ntask) return;
    const int t = P.mt[i], q = P.mq[i], sh = q & 31, peer = q >> 5;
    const size_t s = (size_t)(P.sbase[peer] + P.ms[i]);
    double x[3];
    for (int k = 0; k < 3; k++) {
        const double cs = sh ? P.sbox[6 * s + k] + (double)c_mshift[sh][k] * P.period : P.sbox[6 * s + k];
        x[k] = P.box[6 * (size_t)t + k] - cs;
    }
    double f[4], D[NM], M[NM];
    radial_factors(x[0] * x[0] + x[1] * x[1] + x[2] * x[2], P.rs, f);
    derivative_tensor(x, f, D);
#pragma unroll
    const size_t sm = (size_t)(P.sbase_M[peer] + P.ms[i]);
#pragma unroll
    for (int n = 0; n < NM; n++) M[n] = P.sM[NM * sm + n];
    double* L = P.L + NM * (size_t)t;
#pragma unroll
    for (int n = 0; n < NM; n++) {
        const Idx N = idx_of(n);
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NM; k++) {
            const Idx K = idx_of(k);
            if (N.a + N.b + N.c + K.a + K.b + K.c <= 3) a += M[k] * D[pos_of(N.a + K.a, N.b + K.b, N.c + K.c)];
        }
        atomicAdd(L + n, a);
    }
}

// one thread per node of one tree level, root first: its expansion moves to both sons (walk_l2l)
__global__ void l2l_level_kernel(const int* __restrict__ t_id, int lvl_begin, int lvl_count, const int* __restrict__ son, int nleaf,
                                 const double* __restrict__ box, double* __restrict__ L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lvl_count) return;
    const int id = t_id[lvl_begin + i];
    const size_t u = (size_t)nleaf + id;
    double in[NM];
#pragma unroll
    for (int n = 0; n < NM; n++) in[n] = L[NM * u + n];
    for (int s = 0; s < 2; s++) {
        const int c = son[2 * id + s];
        if (c < 0) continue;
        double pw[NM], out[NM];
        powers(box[6 * (size_t)c] - box[6 * u], box[6 * (size_t)c + 1] - box[6 * u + 1], box[6 * (size_t)c + 2] - box[6 * u + 2], pw);
#pragma unroll
        for (int n = 0; n < NM; n++) out[n] = 0.0;
        shift_down(in, pw, out);
#pragma unroll
        for (int n = 0; n < NM; n++) L[NM * (size_t)c + n] += out[n];
    }
}

// one warp per leaf, one lane per particle: acceleration from the leaf's local expansion (l2p)
__global__ void l2p_kernel(const int2* __restrict__ leaf, int nleaf, const double* __restrict__ box, const double* __restrict__ L,
                           const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                           double* __restrict__ acc_mid) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nleaf) return;
    const int2 lf = leaf[w];
    double F[NM];
#pragma unroll
    for (int n = 0; n < NM; n++) F[n] = L[NM * (size_t)w + n];
    for (int k = lane; k < lf.y; k += 32) {
        const int p = lf.x + k;
        double pw[NM];
        powers(px[p] - box[6 * (size_t)w], py[p] - box[6 * (size_t)w + 1], pz[p] - box[6 * (size_t)w + 2], pw);
        double a[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int n = 0; n < NM; n++) {
            const Idx N = idx_of(n);
            if (N.a + N.b + N.c <= 2) {
                a[0] += F[pos_of(N.a + 1, N.b, N.c)] * pw[n];
                a[1] += F[pos_of(N.a, N.b + 1, N.c)] * pw[n];
                a[2] += F[pos_of(N.a, N.b, N.c + 1)] * pw[n];
            }
        }
        acc_mid[3 * (size_t)p] = a[0];
        acc_mid[3 * (size_t)p + 1] = a[1];
        acc_mid[3 * (size_t)p + 2] = a[2];
    }
}

// acc (P2P, float4, tree order) + mid-field (fp64, tree order) -> caller's order
__global__ void acc_unpermute_sum_kernel(const float4* __restrict__ acc, const double* __restrict__ mid, const int* __restrict__ perm,
                                         long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    double* o = out + 3 * (size_t)perm[i];
    o[0] = (double)a.x + mid[3 * i];
    o[1] = (double)a.y + mid[3 * i + 1];
    o[2] = (double)a.z + mid[3 * i + 2];
}

}  // namespace mf
}  // namespace p2p
