/* rccl.h -- TEST INFRASTRUCTURE ONLY: the reference's RCCL path is compiled out (ENABLE_RCCL is not defined). */
#ifndef REF3D_RCCL_SHIM_H
#define REF3D_RCCL_SHIM_H
typedef int ncclResult_t;
#define ncclSuccess 0
static inline const char* ncclGetErrorString(ncclResult_t) { return "rccl stub"; }
#endif
