// Stub of <mpi.h> for compiling the reference's PARSER sources on one CPU process (oracle/ref, test infrastructure).
#pragma once
#include <cstdio>
#include <cstdlib>
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { fprintf(stderr, "MPI_Abort(%d) from reference code\n", code); abort(); return 0; }
