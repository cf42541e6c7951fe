/* Stub <mpi.h> for building the mounted reference's hot-path host code (fmm.c, remotes.c,
 * toptree.c, domains.c, operator.c) as a single-rank process WITHOUT an MPI installation.
 * TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product library.
 * Only the handful of calls those five files make are declared; see ref_stubs.c for their
 * single-rank semantics (collectives = memcpy, Isend/Recv = self exchange matched by tag). */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0
/* datatype ids double as "bytes per element" where that is unambiguous */
#define MPI_CHAR   1
#define MPI_BYTE   1
#define MPI_INT    4
#define MPI_DOUBLE 8
#define MPI_SUM 1
#define MPI_MAX 2
int MPI_Barrier(MPI_Comm c);
int MPI_Allgather(const void* s, int sc, MPI_Datatype st, void* r, int rc, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoall(const void* s, int sc, MPI_Datatype st, void* r, int rc, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoallv(const void* s, const int* sc, const int* sd, MPI_Datatype st, void* r, const int* rc,
                  const int* rd, MPI_Datatype rt, MPI_Comm c);
int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* rq);
int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st);
int MPI_Wait(MPI_Request* rq, MPI_Status* st);
#endif
