/* hip/hip_runtime.h -- TEST INFRASTRUCTURE ONLY (oracle/ref_3dmpifft).
 * A HIP-on-CPU stand-in, just wide enough to compile the reference's hot-path sources IN PLACE with g++
 * (3dmpifft_opt/include/fft_mpi_3d_api.cpp, kernel_func.cpp, fast_transpose/kernels_201.cpp, kernels_120.cpp) and to run
 * them on host memory: "device" memory is the heap, copies are memcpy, every kernel launch runs its whole grid on the
 * calling thread, one fiber per GPU thread so that __syncthreads() and __shared__ work (hipcpu_launch, ref3d_glue.cpp).
 * The reference drives one GPU per OpenMP thread; blockIdx/threadIdx/... are therefore thread-local.
 * Nothing on the product path includes or links this. */
#ifndef REF3D_HIP_RUNTIME_SHIM_H
#define REF3D_HIP_RUNTIME_SHIM_H
#include <assert.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef __cplusplus
#include <functional>
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct hipcpu_uint3 { unsigned x, y, z; };
extern thread_local hipcpu_uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
void hipcpu_launch(dim3 grid, dim3 block, const std::function<void()>& body);
void hipcpu_syncthreads(void);
#define hipLaunchKernelGGL(kernel, grid, block, shmem, stream, ...) hipcpu_launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define __syncthreads() hipcpu_syncthreads()
#endif

#define __global__
#define __device__
#define __host__
#define __constant__
#define __launch_bounds__(...)
#ifdef HIPCPU_GENERATED_KERNEL
/* a kernel the reference's generator (templateFFT.cpp) emitted at run time, compiled by the hiprtc stand-in (tfft_engine.cpp):
 * it declares `extern __shared__ float shared[];` -- dynamic shared memory, here 64 KB per "device" thread */
#define __shared__ thread_local
thread_local float shared[16384] __attribute__((aligned(16)));
#else
#define __shared__ static thread_local     /* one copy per "device" thread; blocks of a launch run one after the other */
#endif
typedef struct { double x, y; } double2;
typedef struct { float x, y; } float2;

typedef int hipError_t;
#define hipSuccess 0
typedef void* hipStream_t;
typedef int hipDevice_t;
typedef void* hipCtx_t;
typedef void* hipModule_t;
typedef void* hipFunction_t;
typedef void* hipDeviceptr_t;
enum hipDeviceAttribute_t { hipDeviceAttributeMaxThreadsPerBlock, hipDeviceAttributeMaxGridDimX, hipDeviceAttributeMaxGridDimY, hipDeviceAttributeMaxGridDimZ,
                            hipDeviceAttributeMaxBlockDimX, hipDeviceAttributeMaxBlockDimY, hipDeviceAttributeMaxBlockDimZ, hipDeviceAttributeMaxSharedMemoryPerBlock,
                            hipDeviceAttributeWarpSize };
enum hipFuncAttribute { hipFuncAttributeMaxDynamicSharedMemorySize };
enum hipMemcpyKind { hipMemcpyHostToHost = 0, hipMemcpyHostToDevice = 1, hipMemcpyDeviceToHost = 2, hipMemcpyDeviceToDevice = 3, hipMemcpyDefault = 4 };

#ifdef __cplusplus
extern "C" {
#endif
const char* hipGetErrorString(hipError_t e);
hipError_t hipGetDeviceCount(int* n);
hipError_t hipSetDevice(int d);
hipError_t hipDeviceCanAccessPeer(int* can, int a, int b);
hipError_t hipDeviceEnablePeerAccess(int peer, unsigned flags);
hipError_t hipDeviceSynchronize(void);
hipError_t hipFree(void* p);
hipError_t hipMemcpy(void* dst, const void* src, size_t bytes, enum hipMemcpyKind kind);
hipError_t hipMemcpyAsync(void* dst, const void* src, size_t bytes, enum hipMemcpyKind kind, hipStream_t s);
hipError_t hipMemcpyPeerAsync(void* dst, int dstDev, const void* src, int srcDev, size_t bytes, hipStream_t s);
hipError_t hipMemcpyDtoH(void* dst, const void* src, size_t bytes);
hipError_t hipMemcpyHtoD(void* dst, const void* src, size_t bytes);
hipError_t hipStreamCreate(hipStream_t* s);
hipError_t hipStreamDestroy(hipStream_t s);
hipError_t hipStreamSynchronize(hipStream_t s);
hipError_t hipcpu_malloc(void** p, size_t bytes);
/* the module API the reference's FFT engine uses after hiprtc (templateFFT.cpp:5709-5745, 6228): modules are shared objects */
hipError_t hipDeviceGetAttribute(int* value, enum hipDeviceAttribute_t attr, int device);
hipError_t hipModuleLoadDataEx(hipModule_t* module, const void* image, unsigned nopt, void* opts, void* vals);
hipError_t hipModuleUnload(hipModule_t module);
hipError_t hipModuleGetFunction(hipFunction_t* f, hipModule_t module, const char* name);
hipError_t hipModuleGetGlobal(hipDeviceptr_t* ptr, size_t* bytes, hipModule_t module, const char* name);
hipError_t hipFuncSetAttribute(hipFunction_t f, enum hipFuncAttribute attr, int value);
hipError_t hipModuleLaunchKernel(hipFunction_t f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned shmem, hipStream_t s,
                                 void** args, void** extra);
#ifdef __cplusplus
}
template <class T> static inline hipError_t hipMalloc(T** p, size_t bytes) { return hipcpu_malloc((void**)p, bytes); }
#endif
#endif
