/* Stub of <mpi.h> (written for this repo) for compiling reference sources on one CPU process (oracle/ref, test
 * infrastructure).  Valid C and C++. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* r) { (void)c; *r = 1; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d) from reference code\n", code); abort(); return 0; }
typedef int MPI_Op;
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#include <string.h>
#include <time.h>
static inline int MPI_Allreduce(const void* s, void* r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; memcpy(r, s, (size_t)count * (size_t)t); return 0;
}
static inline int MPI_Bcast(void* b, int count, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)count; (void)t; (void)root; (void)c; return 0; }
static inline double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
typedef int MPI_Request;
typedef int MPI_Status;
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_CHAR 1
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)r; return 0; }
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)r; return 0; }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r; (void)s; return 0; }
static inline int MPI_Reduce(const void* s, void* r, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; memcpy(r, s, (size_t)count * (size_t)t); return 0; }
