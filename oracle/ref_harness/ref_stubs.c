/* Single-node, fork-based stand-in for the few MPI calls the reference's hot-path host code
 * makes (see mpi.h in this directory), so that the UNMODIFIED reference sources under
 * /root/reference can run as P >= 1 cooperating "ranks" without an MPI installation.
 * TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product library.
 *
 * Ranks are processes forked by ref_harness.c after stub_mpi_init() has created one shared
 * anonymous mapping that holds: a process-shared barrier, a staging area for the collectives
 * and one byte ring per ordered (src,dst) pair for Isend/Recv (matching is by arrival order
 * within the pair, tag checked).  MPI_Isend copies eagerly into the ring, so MPI_Wait is a no-op.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include "mpi.h"

#define MAXP 16
#define STAGE_BYTES (1u << 20)          /* per-rank staging for collectives */
#define RING_BYTES ((size_t)1 << 31)    /* virtual size per (src,dst) ring; touched lazily */

typedef struct {
    volatile uint64_t head; /* bytes written (producer) */
    volatile uint64_t tail; /* bytes consumed (consumer) */
} RingCtl;

typedef struct {
    pthread_barrier_t bar;
    RingCtl ctl[MAXP][MAXP];
    unsigned char stage[MAXP][STAGE_BYTES];
} Shared;

static Shared* sh;
static unsigned char* rings; /* MAXP*MAXP rings, only [P][P] used */
static int P = 1, R = 0;

static size_t dtsize(MPI_Datatype t) { return t >= 1000 ? (size_t)(t - 1000) : (size_t)t; }

void stub_mpi_init(int nproc) {
    if (nproc > MAXP) { fprintf(stderr, "stub mpi: at most %d ranks\n", MAXP); exit(2); }
    P = nproc;
    sh = (Shared*)mmap(NULL, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    rings = (unsigned char*)mmap(NULL, RING_BYTES * (size_t)P * P, PROT_READ | PROT_WRITE,
                                 MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (sh == MAP_FAILED || rings == MAP_FAILED) { perror("stub mpi mmap"); exit(2); }
    memset((void*)sh->ctl, 0, sizeof(sh->ctl));
    pthread_barrierattr_t a;
    pthread_barrierattr_init(&a);
    pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &a, (unsigned)P);
}
void stub_mpi_set_rank(int r) { R = r; }

int MPI_Barrier(MPI_Comm c) { (void)c; if (P > 1) pthread_barrier_wait(&sh->bar); return 0; }

int MPI_Allgather(const void* s, int sc, MPI_Datatype st, void* r, int rc, MPI_Datatype rt, MPI_Comm c) {
    size_t nb = (size_t)sc * dtsize(st);
    (void)rc; (void)rt;
    if (nb > STAGE_BYTES) { fprintf(stderr, "stub mpi: allgather too large\n"); exit(2); }
    memcpy(sh->stage[R], s, nb);
    MPI_Barrier(c);
    for (int p = 0; p < P; p++) memcpy((char*)r + (size_t)p * nb, sh->stage[p], nb);
    MPI_Barrier(c);
    return 0;
}

int MPI_Alltoall(const void* s, int sc, MPI_Datatype st, void* r, int rc, MPI_Datatype rt, MPI_Comm c) {
    size_t nb = (size_t)sc * dtsize(st);
    (void)rc; (void)rt;
    if (nb * P > STAGE_BYTES) { fprintf(stderr, "stub mpi: alltoall too large\n"); exit(2); }
    memcpy(sh->stage[R], s, nb * P);
    MPI_Barrier(c);
    for (int p = 0; p < P; p++) memcpy((char*)r + (size_t)p * nb, sh->stage[p] + (size_t)R * nb, nb);
    MPI_Barrier(c);
    return 0;
}

int MPI_Alltoallv(const void* s, const int* sc, const int* sd, MPI_Datatype st, void* r, const int* rc,
                  const int* rd, MPI_Datatype rt, MPI_Comm c) {
    (void)s; (void)sc; (void)sd; (void)st; (void)r; (void)rc; (void)rd; (void)rt; (void)c;
    fprintf(stderr, "stub mpi: MPI_Alltoallv unused (reference is built with -DMYALLTOALLV)\n");
    exit(2);
}

static void ring_write(int src, int dst, const void* buf, size_t nb) {
    RingCtl* k = &sh->ctl[src][dst];
    unsigned char* base = rings + ((size_t)src * P + dst) * RING_BYTES;
    const unsigned char* p = (const unsigned char*)buf;
    while (nb) {
        uint64_t h = k->head;
        while (h - k->tail >= RING_BYTES) sched_yield();
        size_t off = (size_t)(h % RING_BYTES);
        size_t room = RING_BYTES - (size_t)(h - k->tail);
        size_t n = nb < room ? nb : room;
        if (n > RING_BYTES - off) n = RING_BYTES - off;
        memcpy(base + off, p, n);
        __sync_synchronize();
        k->head = h + n;
        p += n; nb -= n;
    }
}
static void ring_read(int src, int dst, void* buf, size_t nb) {
    RingCtl* k = &sh->ctl[src][dst];
    unsigned char* base = rings + ((size_t)src * P + dst) * RING_BYTES;
    unsigned char* p = (unsigned char*)buf;
    while (nb) {
        uint64_t t = k->tail;
        while (k->head == t) sched_yield();
        __sync_synchronize();
        size_t avail = (size_t)(k->head - t);
        size_t off = (size_t)(t % RING_BYTES);
        size_t n = nb < avail ? nb : avail;
        if (n > RING_BYTES - off) n = RING_BYTES - off;
        memcpy(p, base + off, n);
        __sync_synchronize();
        k->tail = t + n;
        p += n; nb -= n;
    }
}

/* observers for the harness: sizes of the last messages received with the reference's halo tags */
int stub_last_recv_count[256];

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* rq) {
    (void)c;
    int64_t hdr[2] = {tag, (int64_t)((size_t)count * dtsize(t))};
    ring_write(R, dest, hdr, sizeof hdr);
    ring_write(R, dest, buf, (size_t)hdr[1]);
    if (rq) *rq = 0;
    return 0;
}
int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
    (void)c; (void)st;
    int64_t hdr[2];
    ring_read(src, R, hdr, sizeof hdr);
    size_t want = (size_t)count * dtsize(t);
    if (hdr[0] != tag || (size_t)hdr[1] != want) {
        fprintf(stderr, "stub mpi: rank %d recv from %d: tag %d/%ld bytes %zu/%ld mismatch\n", R, src, tag,
                (long)hdr[0], want, (long)hdr[1]);
        exit(2);
    }
    ring_read(src, R, buf, want);
    if (tag >= 0 && tag < 256 && want == sizeof(int)) stub_last_recv_count[tag] = *(int*)buf;
    return 0;
}
int MPI_Wait(MPI_Request* rq, MPI_Status* st) { (void)rq; (void)st; return 0; }
