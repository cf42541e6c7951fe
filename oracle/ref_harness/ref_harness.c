/* Driver that runs the UNMODIFIED reference host code of the P2P path (1_Indexing: fmm.c,
 * remotes.c, toptree.c, domains.c, operator.c, compiled where they lie under /root/reference by
 * ../Makefile) on a caller-supplied particle set and records what it hands to the GPU C-ABI.
 * TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product library.
 *
 * It mirrors the call sequence of 1_Indexing/src/photoNs.c:83-123 minus PM:
 *   domain_initialize -> [domain_decomposition] -> fmm_construct -> fmm_prepare -> fmm_task -> [fmm_ext]
 * and implements the six symbols of 1_Indexing/inc/photoNs_CUDA.cuh:24-33 as CAPTURE stubs:
 * copyMemGPU records the (target leaf, source leaf) list, readResultsGPU returns zeros.
 *
 * usage: ref_lists <pos.f64> <npart_total> <boxsize> <maxleaf> <nside> <theta> <do_ext 0|1> <nproc> <out_prefix>
 *   pos.f64: npart_total*3 little-endian doubles.  One output file <out_prefix>.rank<r> per rank:
 *   a sequence of records  [u32 name_len][name][u32 dtype: 0=i32 1=f64 2=i64][u64 count][payload].
 */
#define _GNU_SOURCE
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>
#include "photoNs.h"
#include "fmm.h"
#include "domains.h"

void stub_mpi_init(int nproc);
void stub_mpi_set_rank(int r);
extern int stub_last_recv_count[256];

/* ---- replacements for utility.c (the original pmalloc memsets ~3 GB per step) ---- */
double dtime(void) { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }
void* pmalloc(size_t size, int idx) { (void)idx; void* p = calloc(size ? size : 1, 1); if (!p) { fprintf(stderr, "pmalloc(%zu) failed\n", size); exit(3);} return p; }
void pfree(void* p, int idx) { (void)idx; free(p); }
void mem_shift(int s, int t) { (void)s; (void)t; }
void reset_mem(void) {}

/* ---- output container ---- */
static FILE* fout;
static void put(const char* name, int dtype, const void* data, uint64_t count) {
    uint32_t nl = (uint32_t)strlen(name), dt = (uint32_t)dtype;
    size_t es = dtype == 0 ? 4 : 8;
    fwrite(&nl, 4, 1, fout); fwrite(name, 1, nl, fout); fwrite(&dt, 4, 1, fout); fwrite(&count, 8, 1, fout);
    if (count) fwrite(data, es, count, fout);
}
static void put_i(const char* name, int64_t v) { put(name, 2, &v, 1); }
static void put_d(const char* name, double v) { put(name, 1, &v, 1); }

/* ---- capture stubs for the Indexing GPU C-ABI (1_Indexing/inc/photoNs_CUDA.cuh:24-33) ---- */
static int launch_no = 0;            /* 0 = local list (fmm_task); 1.. = fmm_remote calls in order */
static double cur_shift[3];
static double t_walk_local;
void initGPU(int v) { (void)v; }
void getGPUMemoryState(int v) { (void)v; }
int allocMemGPU(int a, int b, int c, int d, int v) { (void)a; (void)b; (void)c; (void)d; (void)v; return 0; }
int copyMemGPU(double* h_pos, int* h_leaf, int* h_int, int ntasks, int v) {
    (void)h_pos; (void)h_leaf; (void)v;
    char nm[64];
    if (launch_no == 0) {
        put("local_tasks_ts", 0, h_int, (uint64_t)ntasks * 2);   /* {target, source} minus first_leaf */
    } else {
        /* remote call: sources index exrtree (h_int[2n+1] + first_leaf); dump the received pruned tree */
        int nnode = stub_last_recv_count[101], nbody = stub_last_recv_count[102];
        snprintf(nm, sizeof nm, "remote%d_shift", launch_no); put(nm, 1, cur_shift, 3);
        snprintf(nm, sizeof nm, "remote%d_tasks_ts", launch_no); put(nm, 0, h_int, (uint64_t)ntasks * 2);
        int* ni = (int*)malloc(sizeof(int) * 3 * (size_t)(nnode ? nnode : 1));
        double* nd = (double*)malloc(sizeof(double) * 6 * (size_t)(nnode ? nnode : 1));
        for (int i = 0; i < nnode; i++) {
            ni[3 * i] = exrtree[i].npart; ni[3 * i + 1] = exrtree[i].son[0]; ni[3 * i + 2] = exrtree[i].son[1];
            for (int k = 0; k < 3; k++) { nd[6 * i + k] = exrtree[i].center[k]; nd[6 * i + 3 + k] = exrtree[i].width[k]; }
        }
        snprintf(nm, sizeof nm, "remote%d_node_npart_son", launch_no); put(nm, 0, ni, (uint64_t)nnode * 3);
        snprintf(nm, sizeof nm, "remote%d_node_center_width", launch_no); put(nm, 1, nd, (uint64_t)nnode * 6);
        double* bd = (double*)malloc(sizeof(double) * 3 * (size_t)(nbody ? nbody : 1));
        for (int i = 0; i < nbody; i++) for (int k = 0; k < 3; k++) bd[3 * i + k] = exrbody[i].pos[k];
        snprintf(nm, sizeof nm, "remote%d_body_pos", launch_no); put(nm, 1, bd, (uint64_t)nbody * 3);
        free(ni); free(nd); free(bd);
    }
    return 0;
}
int LaunchKernelP2PIndexing(int nt, int pc, int lc, int rc, double eps, double m, int v) {
    (void)nt; (void)pc; (void)lc; (void)rc; (void)eps; (void)m; (void)v; return 0;
}
void readResultsGPU(double* h_acc, int nTasks, int maxParts, int v) {
    (void)v; memset(h_acc, 0, sizeof(double) * 3 * (size_t)maxParts * (size_t)nTasks);
    launch_no++;
}

/* fmm_remote is wrapped (ld --wrap) only to learn the displacement of the current call */
void __real_fmm_remote(int idx, double displace[3]);
void __wrap_fmm_remote(int idx, double displace[3]) {
    cur_shift[0] = displace[0]; cur_shift[1] = displace[1]; cur_shift[2] = displace[2];
    int before = launch_no;
    __real_fmm_remote(idx, displace);
    if (launch_no == before) launch_no++;   /* early return (nothing received): keep numbering per call */
}

/* restated from 1_Indexing/src/initial.c:204-228 (initial.c itself drags in the whole program) */
static void setup_domain_index_(void) {
    mostleft = 1;
    while (mostleft < 2 * PROC_SIZE - 1) mostleft *= 2;
    mostleft /= 2; mostleft -= 1;
    if (PROC_SIZE == 1) mostleft = 0;
    this_domain = PROC_RANK + mostleft;
    if (this_domain > 2 * PROC_SIZE - 2) this_domain -= PROC_SIZE;
    first_domain = PROC_SIZE - 1;
    last_domain = 2 * PROC_SIZE - 2;
}

static int run_rank(int rank, int nproc, const double* pos, long ntot, double box, int maxleaf, int nside,
                    double theta, int do_ext, const char* prefix) {
    char fn[512];
    snprintf(fn, sizeof fn, "%s.rank%d", prefix, rank);
    fout = fopen(fn, "wb");
    if (!fout) { perror(fn); return 4; }
    /* silence the reference's own printf noise */
    fflush(stdout); if (!freopen("/dev/null", "w", stdout)) return 4;

    stub_mpi_set_rank(rank);
    PROC_SIZE = nproc; PROC_RANK = rank;
    setup_domain_index_();
    NPART_TOTAL = ntot; BOXSIZE = box; MAXLEAF = maxleaf; NSIDE = nside; open_angle = theta;
    MASSPART = 1.0; verbosity_gpu = 0;
    /* derived exactly as 1_Indexing/src/initial.c:324-346 */
    splitRadius = 1.25 * (BOXSIZE / ((double)NSIDE));
    SoftenScale = 0.03 * BOXSIZE / pow(((double)NPART_TOTAL), 0.3333333);
    cutoffRadius = 4.5 * splitRadius;
    strBody = 1000 + (int)sizeof(Body); strReNode = 1000 + (int)sizeof(RemoteNode); strReBody = 1000 + (int)sizeof(RemoteBody);

    /* initial slab of the global array per rank, original index kept in vel[0] */
    long lo = ntot * rank / nproc, hi = ntot * (rank + 1) / nproc;
    NPART = (int)(hi - lo);
    part = (Body*)calloc((size_t)NPART > 0 ? (size_t)NPART : 1, sizeof(Body));
    for (long i = lo; i < hi; i++) {
        for (int k = 0; k < 3; k++) part[i - lo].pos[k] = pos[3 * i + k];
        part[i - lo].vel[0] = (double)i;
    }

    domain_initialize();
    if (nproc > 1) domain_decomposition();
    fmm_construct();
    double t0 = dtime();
    fmm_prepare();
    double t_build = dtime() - t0;

    put_i("nproc", nproc); put_i("rank", rank); put_i("npart", NPART); put_i("maxleaf", MAXLEAF);
    put_i("first_leaf", first_leaf); put_i("last_leaf", last_leaf); put_i("first_node", first_node); put_i("last_node", last_node);
    put_i("direct_local_start", direct_local_start); put_i("this_domain", this_domain);
    put_d("splitRadius", splitRadius); put_d("cutoffRadius", cutoffRadius); put_d("SoftenScale", SoftenScale);
    {
        int nl = last_leaf - first_leaf, nn = last_node - first_node + 1;
        int64_t* perm = (int64_t*)malloc(sizeof(int64_t) * (size_t)(NPART ? NPART : 1));
        double* pp = (double*)malloc(sizeof(double) * 3 * (size_t)(NPART ? NPART : 1));
        for (int i = 0; i < NPART; i++) { perm[i] = (int64_t)part[i].vel[0]; for (int k = 0; k < 3; k++) pp[3 * i + k] = part[i].pos[k]; }
        put("part_orig_index", 2, perm, (uint64_t)NPART); put("part_pos", 1, pp, (uint64_t)NPART * 3);
        int* li = (int*)malloc(sizeof(int) * 2 * (size_t)(nl ? nl : 1));
        double* ld = (double*)malloc(sizeof(double) * 6 * (size_t)(nl ? nl : 1));
        for (int i = 0; i < nl; i++) {
            Pack* l = &leaf[first_leaf + i];
            li[2 * i] = l->npart; li[2 * i + 1] = l->ipart;
            for (int k = 0; k < 3; k++) { ld[6 * i + k] = l->center[k]; ld[6 * i + 3 + k] = l->width[k]; }
        }
        put("leaf_npart_ipart", 0, li, (uint64_t)nl * 2); put("leaf_center_width", 1, ld, (uint64_t)nl * 6);
        int* ni = (int*)malloc(sizeof(int) * 3 * (size_t)nn);
        double* nd = (double*)malloc(sizeof(double) * 7 * (size_t)nn);
        for (int i = 0; i < nn; i++) {
            Node* b = &btree[first_node + i];
            ni[3 * i] = b->npart; ni[3 * i + 1] = b->son[0]; ni[3 * i + 2] = b->son[1];
            for (int k = 0; k < 3; k++) { nd[7 * i + k] = b->center[k]; nd[7 * i + 3 + k] = b->width[k]; }
            nd[7 * i + 6] = b->split;
        }
        put("node_npart_son", 0, ni, (uint64_t)nn * 3); put("node_center_width_split", 1, nd, (uint64_t)nn * 7);
        double* td = (double*)malloc(sizeof(double) * 7 * (size_t)(2 * nproc - 1));
        for (int i = 0; i < 2 * nproc - 1; i++) {
            for (int k = 0; k < 3; k++) { td[7 * i + k] = toptree[i].center[k]; td[7 * i + 3 + k] = toptree[i].width[k]; }
            td[7 * i + 6] = toptree[i].split;
        }
        put("toptree_center_width_split", 1, td, (uint64_t)(2 * nproc - 1) * 7);
        free(perm); free(pp); free(li); free(ld); free(ni); free(nd); free(td);
    }
    t0 = dtime();
    fmm_task();
    t_walk_local = dtime() - t0;
    if (launch_no == 0) launch_no = 1;
    put_d("t_build_s", t_build); put_d("t_fmm_task_s", t_walk_local);
    put_i("idxP2P_local", (int64_t)idxP2P);
    put_i("idxM2L_local", (int64_t)idxM2L);
    {   /* multipoles after fmm_prepare (p2m + walk_m2m): the mid-field oracle of row N2 */
        int nl = last_leaf - first_leaf, nn = last_node - first_node + 1;
        double* m = (double*)malloc(sizeof(double) * NMULTI * (size_t)((nl > nn ? nl : nn) + 1));
        for (int i = 0; i < nl; i++) memcpy(m + (size_t)NMULTI * i, leaf[first_leaf + i].M, sizeof(double) * NMULTI);
        put("leaf_M", 1, m, (uint64_t)nl * NMULTI);
        for (int i = 0; i < nn; i++) memcpy(m + (size_t)NMULTI * i, btree[first_node + i].M, sizeof(double) * NMULTI);
        put("node_M", 1, m, (uint64_t)nn * NMULTI);
        free(m);
    }
    if (do_ext) {
        t0 = dtime();
        fmm_ext();
        put_d("t_fmm_ext_s", dtime() - t0);
        put_i("idxM2L_total", (int64_t)idxM2L);
        {   /* after walk_l2l + l2p: local expansions of the leaves and the mid-field accelerations (the P2P stubs above
             * return zeros, so part[].acc holds M2L -> L2L -> L2P only), tree order */
            int nl = last_leaf - first_leaf;
            double* m = (double*)malloc(sizeof(double) * NMULTI * (size_t)(nl + 1));
            for (int i = 0; i < nl; i++) memcpy(m + (size_t)NMULTI * i, leaf[first_leaf + i].L, sizeof(double) * NMULTI);
            put("leaf_L", 1, m, (uint64_t)nl * NMULTI);
            free(m);
            double* a = (double*)malloc(sizeof(double) * 3 * (size_t)(NPART ? NPART : 1));
            for (int i = 0; i < NPART; i++) for (int k = 0; k < 3; k++) a[3 * i + k] = part[i].acc[k];
            put("acc_mid", 1, a, (uint64_t)NPART * 3);
            free(a);
        }
        put_i("numRemoteInteractions", numRemoteInteractions);
        put_i("n_remote_calls", launch_no - 1);
        /* Second decomposition with the reference's load-balance feedback (1_Indexing/src/photoNs.c:295-306:
         * DTIME_FRACTION = W_r * P / (sum W + 1e-4), W = idxP2P + idxM2L set at 1_Indexing/src/fmm.c:1139),
         * then domain_decomposition() as the next step of driver() would call it (photoNs.c:213). */
        put_d("work_this_domain", DTIME_THIS_DOMAIN);
        {
            double* all = (double*)malloc(sizeof(double) * (size_t)nproc);
            MPI_Allgather(&DTIME_THIS_DOMAIN, 1, MPI_DOUBLE, all, 1, MPI_DOUBLE, MPI_COMM_WORLD);
            double tot = 0.0;
            for (int r = 0; r < nproc; r++) tot += all[r];
            DTIME_FRACTION = DTIME_THIS_DOMAIN * PROC_SIZE / (tot + 0.0001);
            free(all);
        }
        put_d("dtime_fraction", DTIME_FRACTION);
        fmm_deconstruct();
        domain_decomposition();
        {
            double* sp = (double*)malloc(sizeof(double) * (size_t)(2 * nproc - 1));
            for (int i = 0; i < 2 * nproc - 1; i++) sp[i] = domtree[i].split;
            put("domtree_split_step2", 1, sp, (uint64_t)(2 * nproc - 1));
            free(sp);
            int64_t* perm = (int64_t*)malloc(sizeof(int64_t) * (size_t)(NPART ? NPART : 1));
            for (int i = 0; i < NPART; i++) perm[i] = (int64_t)part[i].vel[0];
            put_i("npart_step2", NPART);
            put("part_orig_index_step2", 2, perm, (uint64_t)NPART);
            free(perm);
        }
    }
    fclose(fout);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 10) {
        fprintf(stderr, "usage: %s pos.f64 npart box maxleaf nside theta do_ext nproc out_prefix\n", argv[0]);
        return 1;
    }
    long ntot = atol(argv[2]);
    double box = atof(argv[3]); int maxleaf = atoi(argv[4]), nside = atoi(argv[5]);
    double theta = atof(argv[6]); int do_ext = atoi(argv[7]), nproc = atoi(argv[8]);
    double* pos = (double*)malloc(sizeof(double) * 3 * (size_t)ntot);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(pos, sizeof(double), 3 * (size_t)ntot, f) != 3 * (size_t)ntot) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    stub_mpi_init(nproc);
    if (nproc == 1) return run_rank(0, 1, pos, ntot, box, maxleaf, nside, theta, do_ext, argv[9]);
    pid_t pids[64];
    for (int r = 0; r < nproc; r++) {
        pids[r] = fork();
        if (pids[r] == 0) _exit(run_rank(r, nproc, pos, ntot, box, maxleaf, nside, theta, do_ext, argv[9]));
    }
    int rc = 0, st;
    while (wait(&st) > 0) {
        if (!WIFEXITED(st) || WEXITSTATUS(st)) {
            /* a rank died (the reference overruns its fixed halo buffers on very unbalanced inputs):
             * its peers would wait for it forever, so stop them */
            rc = 5;
            for (int r = 0; r < nproc; r++) kill(pids[r], SIGKILL);
        }
    }
    return rc;
}
