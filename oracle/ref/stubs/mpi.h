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
