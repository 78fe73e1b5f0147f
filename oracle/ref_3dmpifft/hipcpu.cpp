// hipcpu.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_3dmpifft -> oracle/_ref/libhipcpu.so).
// The HIP-on-CPU runtime behind hip/hip_runtime.h of this directory: "device" memory is the heap, copies are memcpy, and a
// kernel launch runs its whole grid on the calling thread, block by block, with ONE ucontext FIBER PER GPU THREAD -- every sweep
// over the live fibers of a block runs each of them up to its next __syncthreads() (or its end), which is exactly a block
// barrier, so __shared__ tiles and __syncthreads() behave as on the device.  The reference drives one GPU per OpenMP thread;
// blockIdx / threadIdx / ... and the fibers are therefore thread-local.  Shared by the compiled reference sources
// (libref3dmpifft.so), the reference's FFT engine (libtemplatefft_cpu.so) and the kernels that engine generates at run time.
#include <ucontext.h>

#include <vector>

#include "hip/hip_runtime.h"

thread_local hipcpu_uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
static int g_devices = 8;

extern "C" {
void hipcpu_set_device_count(int n) { g_devices = n; }
const char* hipGetErrorString(hipError_t e) { return e ? "hip-on-cpu error" : "no error"; }
hipError_t hipGetDeviceCount(int* n) { *n = g_devices; return hipSuccess; }
hipError_t hipSetDevice(int) { return hipSuccess; }
hipError_t hipDeviceCanAccessPeer(int* can, int, int) { *can = 1; return hipSuccess; }
hipError_t hipDeviceEnablePeerAccess(int, unsigned) { return hipSuccess; }
hipError_t hipDeviceSynchronize(void) { return hipSuccess; }
hipError_t hipcpu_malloc(void** p, size_t bytes) { *p = calloc(1, bytes ? bytes : 1); return *p ? hipSuccess : 2; }
hipError_t hipFree(void* p) { free(p); return hipSuccess; }
hipError_t hipMemcpy(void* d, const void* s, size_t n, enum hipMemcpyKind) { if (d != s) memmove(d, s, n); return hipSuccess; }
hipError_t hipMemcpyAsync(void* d, const void* s, size_t n, enum hipMemcpyKind, hipStream_t) { if (d != s) memmove(d, s, n); return hipSuccess; }
hipError_t hipMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, hipStream_t) { if (d != s) memmove(d, s, n); return hipSuccess; }
hipError_t hipMemcpyDtoH(void* d, const void* s, size_t n) { memmove(d, s, n); return hipSuccess; }
hipError_t hipMemcpyHtoD(void* d, const void* s, size_t n) { memmove(d, s, n); return hipSuccess; }
hipError_t hipStreamCreate(hipStream_t* s) { *s = nullptr; return hipSuccess; }
hipError_t hipStreamDestroy(hipStream_t) { return hipSuccess; }
hipError_t hipStreamSynchronize(hipStream_t) { return hipSuccess; }
// the AMD part the reference targets: 64-wide wavefronts, 64 KB of LDS per work-group
hipError_t hipDeviceGetAttribute(int* v, enum hipDeviceAttribute_t a, int)
{
    switch (a) {
    case hipDeviceAttributeMaxThreadsPerBlock: *v = 1024; break;
    case hipDeviceAttributeMaxGridDimX: *v = 2147483647; break;
    case hipDeviceAttributeMaxGridDimY: case hipDeviceAttributeMaxGridDimZ: *v = 65535; break;
    case hipDeviceAttributeMaxBlockDimX: case hipDeviceAttributeMaxBlockDimY: case hipDeviceAttributeMaxBlockDimZ: *v = 1024; break;
    case hipDeviceAttributeMaxSharedMemoryPerBlock: *v = 65536; break;
    case hipDeviceAttributeWarpSize: *v = 64; break;
    default: return 1;
    }
    return hipSuccess;
}
}

namespace {
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done;
};
thread_local ucontext_t t_sched;
thread_local Fiber* t_cur = nullptr;
thread_local const std::function<void()>* t_body = nullptr;
thread_local std::vector<Fiber> t_fibers;

void fiber_main()
{
    (*t_body)();
    t_cur->done = true;
    swapcontext(&t_cur->ctx, &t_sched);
}
}  // namespace

void hipcpu_syncthreads(void) { swapcontext(&t_cur->ctx, &t_sched); }

void hipcpu_launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    const size_t nt = (size_t)block.x * block.y * block.z;
    if (t_fibers.size() < nt) t_fibers.resize(nt);
    gridDim = grid;
    blockDim = block;
    t_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx = {bx, by, bz};
                for (size_t t = 0; t < nt; t++) {
                    Fiber& f = t_fibers[t];
                    if (f.stack.empty()) f.stack.resize(64 << 10);
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack.data();
                    f.ctx.uc_stack.ss_size = f.stack.size();
                    f.ctx.uc_link = &t_sched;
                    makecontext(&f.ctx, fiber_main, 0);
                    f.done = false;
                }
                for (size_t live = nt; live;)          // one sweep = one block barrier
                    for (size_t t = 0; t < nt; t++) {
                        Fiber& f = t_fibers[t];
                        if (f.done) continue;
                        threadIdx = {(unsigned)(t % block.x), (unsigned)(t / block.x % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                        t_cur = &f;
                        swapcontext(&t_sched, &f.ctx);
                        if (f.done) live--;
                    }
            }
}
