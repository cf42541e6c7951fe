/* Drop-in driver: the UNMODIFIED reference host code (1_Indexing: fmm.c, remotes.c, toptree.c,
 * domains.c, operator.c, compiled where they lie under /root/reference) LINKED AGAINST OUR
 * libphotoNs_CUDA_indexing.so instead of the reference's photoNs_CUDA.cu.
 * TEST INFRASTRUCTURE ONLY (oracle/): built by ../Makefile into _ref/ref_dropin, run by
 * tests/test_gpu_compat.py on the GPU box.
 *
 * It runs domain_initialize -> fmm_construct -> fmm_prepare -> fmm_task (single rank) exactly as
 * 1_Indexing/src/photoNs.c:83-109 does, i.e. the reference's own task_compute_p2p
 * (1_Indexing/src/fmm.c:842-911) packs the buffers, calls initGPU / allocMemGPU / copyMemGPU /
 * LaunchKernelP2PIndexing / readResultsGPU in OUR library and reduces the result slots into
 * part[].acc; the accelerations are written out in the ORIGINAL particle order.
 *
 * usage: ref_dropin <pos.f64> <npart> <boxsize> <maxleaf> <nside> <theta> <mass> <out.f64>
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "photoNs.h"
#include "fmm.h"
#include "domains.h"

void stub_mpi_init(int nproc);
void stub_mpi_set_rank(int r);

double dtime(void) { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }
void* pmalloc(size_t size, int idx) { (void)idx; void* p = calloc(size ? size : 1, 1); if (!p) { fprintf(stderr, "pmalloc(%zu) failed\n", size); exit(3);} return p; }
void pfree(void* p, int idx) { (void)idx; free(p); }
void mem_shift(int s, int t) { (void)s; (void)t; }
void reset_mem(void) {}
/* fmm_remote is only reached from fmm_ext, which this driver does not call */

/* wall time inside the three blocking GPU calls of task_compute_p2p (linked with --wrap, see ../Makefile) */
static double t_copy, t_launch, t_read;
int __real_copyMemGPU(double*, int*, int*, int, int);
int __real_LaunchKernelP2PIndexing(int, int, int, int, double, double, int);
void __real_readResultsGPU(double*, int, int, int);
int __wrap_copyMemGPU(double* a, int* b, int* c, int n, int v) { double t = dtime(); int r = __real_copyMemGPU(a, b, c, n, v); t_copy += dtime() - t; return r; }
int __wrap_LaunchKernelP2PIndexing(int n, int a, int b, int c, double e, double m, int v) {
    double t = dtime(); int r = __real_LaunchKernelP2PIndexing(n, a, b, c, e, m, v); t_launch += dtime() - t; return r; }
void __wrap_readResultsGPU(double* h, int n, int m, int v) { double t = dtime(); __real_readResultsGPU(h, n, m, v); t_read += dtime() - t; }

int main(int argc, char** argv) {
    if (argc != 9) { fprintf(stderr, "usage: %s pos.f64 npart box maxleaf nside theta mass out.f64\n", argv[0]); return 1; }
    long ntot = atol(argv[2]);
    double* pos = (double*)malloc(sizeof(double) * 3 * (size_t)ntot);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(pos, sizeof(double), 3 * (size_t)ntot, f) != 3 * (size_t)ntot) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    stub_mpi_init(1); stub_mpi_set_rank(0);
    PROC_SIZE = 1; PROC_RANK = 0; mostleft = 0; this_domain = first_domain = last_domain = 0;
    NPART_TOTAL = ntot; NPART = (int)ntot; BOXSIZE = atof(argv[3]); MAXLEAF = atoi(argv[4]); NSIDE = atoi(argv[5]);
    open_angle = atof(argv[6]); MASSPART = atof(argv[7]); verbosity_gpu = 0;
    splitRadius = 1.25 * (BOXSIZE / ((double)NSIDE));                       /* 1_Indexing/src/initial.c:324-346 */
    SoftenScale = 0.03 * BOXSIZE / pow(((double)NPART_TOTAL), 0.3333333);
    cutoffRadius = 4.5 * splitRadius;
    part = (Body*)calloc((size_t)NPART, sizeof(Body));
    for (long i = 0; i < ntot; i++) { for (int k = 0; k < 3; k++) part[i].pos[k] = pos[3 * i + k]; part[i].vel[0] = (double)i; }
    int so = dup(1);
    fflush(stdout); if (!freopen("/dev/null", "w", stdout)) return 4;     /* the reference prints a lot */
    domain_initialize();
    fmm_construct();
    fmm_prepare();
    double t_task = dtime();
    fmm_task();
    t_task = dtime() - t_task;
    /* a second pass: device buffers, context and module are warm now (the reference allocates once) */
    for (int i = 0; i < NPART; i++) for (int k = 0; k < 3; k++) part[i].acc[k] = 0.0;
    t_copy = t_launch = t_read = 0.0;
    double t_task2 = dtime();
    fmm_task();
    t_task2 = dtime() - t_task2;
    fprintf(stderr, "ref_dropin: second pass GPU calls: copyMemGPU %.6f s, LaunchKernelP2PIndexing (synchronous) %.6f s, readResultsGPU %.6f s\n",
            t_copy, t_launch, t_read);
    fprintf(stderr, "ref_dropin: fmm_task first %.6f s, second %.6f s (walk + pack + H2D + kernel + D2H + host reduction)\n", t_task, t_task2);
    (void)so;
    double* acc = (double*)malloc(sizeof(double) * 3 * (size_t)ntot);
    for (int i = 0; i < NPART; i++) { long o = (long)part[i].vel[0]; for (int k = 0; k < 3; k++) acc[3 * o + k] = part[i].acc[k]; }
    f = fopen(argv[8], "wb");
    if (!f || fwrite(acc, sizeof(double), 3 * (size_t)ntot, f) != 3 * (size_t)ntot) return 5;
    fclose(f);
    fprintf(stderr, "ref_dropin: %ld particles, idxP2P = %lu tasks through libphotoNs_CUDA_indexing.so\n", ntot, (unsigned long)idxP2P);
    return 0;
}
