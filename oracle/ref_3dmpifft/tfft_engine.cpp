// tfft_engine.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_3dmpifft -> oracle/_ref/libtemplatefft_cpu.so).
//
// The reference's FFT ENGINE run on the CPU.  3dmpifft_opt links a prebuilt libtemplatefft.so (HIP); its source in the tree is
// templateFFT/src/templateFFT.cpp -- a generator that writes one HIP kernel per axis as text (shaderGenFFT), compiles it with
// hiprtc and launches it with hipModuleLaunchKernel.  That file is compiled here in place, unmodified, and this file gives it
//   * hiprtc: hiprtcCompileProgram writes the generated source to oracle/_ref/jit/<hash>.cpp and runs
//     `g++ -shared -include hip/hip_runtime.h -DHIPCPU_GENERATED_KERNEL` on it (cached by source hash); the "code" is the path;
//   * the module API: hipModuleLoadDataEx = dlopen, hipModuleGetFunction / hipModuleGetGlobal = dlsym ("FFT_main", "consts"),
//     hipModuleLaunchKernel = hipcpu_launch of FFT_main(inputs, outputs, twiddleLUT) with one fiber per GPU thread;
// so the butterflies, twiddle LUTs, shared-memory exchanges and index arithmetic that run are the ones the reference's
// generator emitted.  The C entry points below are what ref3d_glue.cpp's initializeFFT / launchFFTKernel forward to (the two
// trees' templateFFT.h differ in field names, so the two translation units cannot share the structs): they fill the
// configuration exactly as fft_mpi_3d_api.cpp:379-431 does.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <mutex>
#include <string>

#include "templateFFT.h"          // templateFFT/src/templateFFT.h (through -I$(REF)/templateFFT/src)
#include <hip/hiprtc.h>

struct hiprtc_program_s {
    std::string src, so, log;
};

static std::string lib_dir()
{
    Dl_info info;
    if (!dladdr((void*)&lib_dir, &info) || !info.dli_fname) return ".";
    std::string p(info.dli_fname);
    const size_t k = p.rfind('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}

extern "C" {
const char* hiprtcGetErrorString(enum hiprtcResult r) { return r == HIPRTC_SUCCESS ? "success" : "g++ failed on the generated kernel"; }
enum hiprtcResult hiprtcCreateProgram(hiprtcProgram* prog, const char* src, const char*, int, const char**, const char**)
{
    *prog = new hiprtc_program_s{src, "", ""};
    return HIPRTC_SUCCESS;
}
enum hiprtcResult hiprtcAddNameExpression(hiprtcProgram, const char*) { return HIPRTC_SUCCESS; }
enum hiprtcResult hiprtcCompileProgram(hiprtcProgram p, int, const char**)
{
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char c : p->src) h = (h ^ c) * 1099511628211ull;
    const std::string dir = lib_dir(), jit = dir + "/jit";
    mkdir(jit.c_str(), 0755);
    char name[64];
    snprintf(name, sizeof(name), "/k%016llx", h);
    const std::string base = jit + name, so = base + ".so";
    if (access(so.c_str(), R_OK) != 0) {
        const std::string cpp = base + ".cpp", tmp = base + "." + std::to_string((long)getpid()) + ".tmp.so";
        FILE* f = fopen(cpp.c_str(), "w");
        if (!f) { p->log = "cannot write " + cpp; return HIPRTC_ERROR_COMPILATION; }
        fputs(p->src.c_str(), f);
        fclose(f);
        // the shim header lives next to this file's sources: <repo>/oracle/ref_3dmpifft (dir = <repo>/oracle/_ref)
        const std::string cmd = "/usr/bin/g++ -O1 -std=gnu++17 -fPIC -shared -w -DHIPCPU_GENERATED_KERNEL -I" + dir + "/../ref_3dmpifft -include hip/hip_runtime.h -o " +
                                tmp + " " + cpp + " " + dir + "/libhipcpu.so -Wl,-rpath," + dir + " 2> " + base + ".log";
        if (system(cmd.c_str()) != 0) { p->log = "failed: " + cmd; return HIPRTC_ERROR_COMPILATION; }
        rename(tmp.c_str(), so.c_str());
    }
    p->so = so;
    return HIPRTC_SUCCESS;
}
enum hiprtcResult hiprtcGetProgramLog(hiprtcProgram p, char* log) { strcpy(log, p->log.c_str()); return HIPRTC_SUCCESS; }
enum hiprtcResult hiprtcGetCodeSize(hiprtcProgram p, size_t* n) { *n = p->so.size() + 1; return HIPRTC_SUCCESS; }
enum hiprtcResult hiprtcGetCode(hiprtcProgram p, char* code) { memcpy(code, p->so.c_str(), p->so.size() + 1); return HIPRTC_SUCCESS; }
enum hiprtcResult hiprtcDestroyProgram(hiprtcProgram* p) { delete *p; *p = nullptr; return HIPRTC_SUCCESS; }

hipError_t hipModuleLoadDataEx(hipModule_t* m, const void* image, unsigned, void*, void*)
{
    *m = dlopen((const char*)image, RTLD_NOW | RTLD_LOCAL);
    if (!*m) fprintf(stderr, "hipModuleLoadDataEx: %s\n", dlerror());
    return *m ? hipSuccess : 1;
}
hipError_t hipModuleUnload(hipModule_t m) { if (m) dlclose(m); return hipSuccess; }
hipError_t hipModuleGetFunction(hipFunction_t* f, hipModule_t m, const char* name) { *f = dlsym(m, name); return *f ? hipSuccess : 1; }
hipError_t hipModuleGetGlobal(hipDeviceptr_t* p, size_t*, hipModule_t m, const char* name) { *p = dlsym(m, name); return *p ? hipSuccess : 1; }
hipError_t hipFuncSetAttribute(hipFunction_t, enum hipFuncAttribute, int bytes) { return bytes <= 65536 ? hipSuccess : 1; }
hipError_t hipModuleLaunchKernel(hipFunction_t f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned shmem, hipStream_t,
                                 void** args, void**)
{
    if (shmem > 65536) return 1;
    typedef void (*kernel_t)(double2*, double2*, double2*);       // FFT_main(inputs, outputs, twiddleLUT), templateFFT.cpp:4821-4830
    double2* a0 = *(double2**)args[0];
    double2* a1 = *(double2**)args[1];
    double2* a2 = *(double2**)args[2];
    hipcpu_launch(dim3(gx, gy, gz), dim3(bx, by, bz), [&]() { ((kernel_t)f)(a0, a1, a2); });
    return hipSuccess;
}

// ---- what ref3d_glue.cpp calls --------------------------------------------------------------------------------------------
struct tfft_handle {
    FFTApplication app;
    FFTConfiguration cfg;
    GPU gpu;
    FFTLaunchParams params;
    uint64_t buffer_size;
    void* dummy;
};
static std::mutex g_mu;      // the generator is not known to be re-entrant; plans are created one at a time

static bool smooth7(long long n)
{
    if (n < 1) return false;
    for (int p : {2, 3, 5, 7}) while (n % p == 0) n /= p;
    return n == 1;
}

void* tfft_create(int fftdim, long long s0, long long s1, long long s2, int inverse)
{
    // the generator's radices are 2, 3, 4, 5, 7, 8 (templateFFT.cpp:331-869); it returns an error for a prime length like 11 but
    // divides by zero (SIGFPE) on composites with a larger prime factor such as 22, so those never reach it
    if (fftdim < 1 || fftdim > 2 || !smooth7(s0) || (fftdim == 2 && !smooth7(s1))) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    tfft_handle* h = new tfft_handle();
    memset(&h->app, 0, sizeof(h->app)); memset(&h->cfg, 0, sizeof(h->cfg)); memset(&h->gpu, 0, sizeof(h->gpu)); memset(&h->params, 0, sizeof(h->params));
    h->cfg.FFTdim = (uint64_t)fftdim;                     // api.cpp:381-431
    h->cfg.size[0] = (uint64_t)s0; h->cfg.size[1] = (uint64_t)s1; h->cfg.size[2] = (uint64_t)s2;
    h->cfg.doublePrecision = true;
    h->cfg.useLUT = true;
    h->cfg.device = &h->gpu.device;
    h->buffer_size = (uint64_t)s0 * (uint64_t)(s1 ? s1 : 1) * (uint64_t)(s2 ? s2 : 1) * 16;
    h->cfg.bufferSize = &h->buffer_size;
    if (inverse) h->cfg.makeInversePlanOnly = true; else h->cfg.makeForwardPlanOnly = true;
    h->dummy = nullptr;
    h->cfg.buffer = &h->dummy;
    if (initializeFFT(&h->app, h->cfg) != FFT_SUCCESS) { delete h; return nullptr; }
    if (setFFTArgs(&h->gpu, &h->app, &h->params, inverse) != FFT_SUCCESS) { delete h; return nullptr; }
    return h;
}
int tfft_launch(void* handle, void** buffer, int inverse)
{
    tfft_handle* h = (tfft_handle*)handle;
    h->app.configuration.buffer = buffer;                 // api.cpp:502 / :517: the buffer is re-pointed before every launch
    return (int)launchFFTKernel(&h->app, inverse);
}
void tfft_destroy(void* handle);
// the radix schedule the generator chose for a 1-D transform of n points: radices of upload 0, then of upload 1, ... (one upload
// = one kernel; more than one = the four-step path for lengths beyond a shared-memory line); *uploads gets their number
int tfft_schedule(long long n, int* radices, int max_radices, int* uploads)
{
    tfft_handle* h = (tfft_handle*)tfft_create(1, n, 1, 1, 0);
    if (!h) return -1;
    FFTPlan* plan = h->app.localFFTPlan;
    int k = 0;
    *uploads = (int)plan->numAxisUploads[0];
    for (uint64_t u = 0; u < plan->numAxisUploads[0]; u++)
        for (uint64_t i = 0; i < plan->axes[0][u].layout.numStages && k < max_radices; i++) radices[k++] = (int)plan->axes[0][u].layout.stageRadix[i];
    tfft_destroy(h);
    return k;
}

void tfft_destroy(void* handle)
{
    if (!handle) return;
    std::lock_guard<std::mutex> lk(g_mu);
    tfft_handle* h = (tfft_handle*)handle;
    deleteFFT(&h->app);
    delete h;
}
}
