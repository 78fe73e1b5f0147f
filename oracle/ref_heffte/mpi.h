/* mpi.h -- TEST INFRASTRUCTURE ONLY (oracle/): a thread-backed stand-in for the handful of MPI calls the
 * vendored heFFTe 2.1.0 uses (heffte/heffteBenchmark/include/heffte_utils.h:75-175, heffte_geometry.h:646,
 * src/heffte_reshape3d.cpp:268, 375, 497-625).  There is no MPI in this image; the "ranks" of a job are the
 * threads started by tmpi_run().  Lets the reference's own CPU FFT (heFFTe + its `stock` backend, the
 * library the reference benchmarks against, heffte/heffteBenchmark/benchmarks/speed3d.h) be compiled from
 * the sources where they lie under /root/reference into oracle/_ref/libheffte_ref.so.
 * Nothing on the product path includes this file. */
#ifndef DFFT_ORACLE_TMPI_H
#define DFFT_ORACLE_TMPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tmpi_comm* MPI_Comm;
typedef struct tmpi_group* MPI_Group;
typedef struct tmpi_req* MPI_Request;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

MPI_Comm tmpi_world(void);
#define MPI_COMM_WORLD (tmpi_world())
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)

/* datatype handle = element size in bytes in the low byte, a distinguishing id above it */
#define MPI_BYTE 0x101
#define MPI_INT 0x204
#define MPI_FLOAT 0x304
#define MPI_DOUBLE 0x408
#define MPI_C_COMPLEX 0x508
#define MPI_C_DOUBLE_COMPLEX 0x610

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_group(MPI_Comm comm, MPI_Group* group);
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group* newgroup);
int MPI_Group_free(MPI_Group* group);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int sendcounts[], const int sdispls[], MPI_Datatype sendtype, void* recvbuf,
                  const int recvcounts[], const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Waitany(int count, MPI_Request reqs[], int* index, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status* statuses);

/* launcher: runs fn(arg) on `nranks` threads that form MPI_COMM_WORLD; returns when all have returned */
int tmpi_run(int nranks, void (*fn)(void*), void* arg);
/* optional pinning: rank i runs on CPU cpus[i % n] (pass one CPU per physical core); n = 0 clears it */
int tmpi_set_cpus(const int* cpus, int n);

#ifdef __cplusplus
}
#endif
#endif
