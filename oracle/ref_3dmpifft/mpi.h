/* mpi.h -- TEST INFRASTRUCTURE ONLY: a one-rank MPI for the reference's single-process mode (all GPUs of the node driven by the
 * OpenMP threads of one rank -- the mode its speedTest.sh runs per node, fftSpeed3d_c2c.cpp:49).  With one rank
 * slabAlltoall (fft_mpi_3d_api.cpp:610-672) moves every block with hipMemcpyPeerAsync and posts no MPI request. */
#ifndef REF3D_MPI_SHIM_H
#define REF3D_MPI_SHIM_H
#include <time.h>
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef struct { int unused; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 1
#define MPI_BYTE 2
#define MPI_MAX 3
#define MPI_THREAD_SERIALIZED 2
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline double MPI_Wtime(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static inline int MPI_Init_thread(int*, char***, int required, int* provided) { *provided = required; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, int, int, MPI_Comm) { __builtin_memcpy(r, s, (t == MPI_DOUBLE ? 8 : 1) * (unsigned long)n); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return 1; }   /* never reached with one rank */
static inline int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return 1; }
static inline int MPI_Waitall(int n, MPI_Request*, MPI_Status*) { return n == 0 ? MPI_SUCCESS : 1; }
#endif
