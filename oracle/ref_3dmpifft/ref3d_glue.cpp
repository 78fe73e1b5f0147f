// ref3d_glue.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_3dmpifft -> oracle/_ref/libref3dmpifft.so).
//
// Runs the REFERENCE'S OWN hot-path sources on the CPU.  Compiled from /root/reference, in place, with g++:
//     3dmpifft_opt/include/fft_mpi_3d_api.cpp   plan creation (TransInfo tables, buffer roles), fftZY, localTransposeUneven,
//                                               slabAlltoall, fftX, fft_mpi_execute_dft_3d_c2c, the count helpers
//     3dmpifft_opt/include/kernel_func.cpp      the t1 pack / unpack kernels (hipLaunchKernelGGL launches)
//     3dmpifft_opt/include/fast_transpose/kernels_201.cpp, kernels_120.cpp   the cuTranspose tile kernels of fftX
// against the HIP-on-CPU headers of this directory (runtime: hipcpu.cpp -- heap memory, memcpy, kernel launches on fibers so
// that __shared__ tiles and __syncthreads() behave as on the device).  What this file supplies:
//   * the four templateFFT entry points the reference calls (templateFFT.h:361-365).  Engine 1 (default): they forward to
//     libtemplatefft_cpu.so = the reference's own FFT engine (templateFFT/src/templateFFT.cpp, compiled in place) whose
//     run-time-generated kernels are compiled with g++ and run on fibers (tfft_engine.cpp) -- the reference's butterflies and
//     twiddles.  Engine 0: a plain DFT with the engine's semantics: an unnormalised transform over the first FFTdim axes of
//     `size` (axis 0 fastest), in place on *configuration.buffer as it is at launch time, every line of the remaining axes
//     being a batch (templateFFT/src/templateFFT.cpp:6073-6095 axis 0, :6106-6109 axis 1).  The two must agree (tested);
//   * cut_transpose3d: fast_transpose/transpose3d.cpp launches with <<< >>> and cannot go through g++, so its dispatcher is
//     restated for the two out-of-place permutations fftX uses (transpose3d.cpp:198-224 -> 120, :225-262 -> 201, grid from
//     set_grid_dims :312-330); the kernels it launches are the reference's;
//   * ref3d_run(): the call sequence of the reference driver (fftSpeed3d_c2c.cpp:42-102), one OpenMP thread per device.
// So with engine 1 everything that computes or moves data -- slab bookkeeping, exchange tables, FFT kernels, pack / unpack
// maps, all-to-all offsets, transposes, stage order -- is the reference's code, executed.
#include <dlfcn.h>
#include <omp.h>

#include <complex>
#include <string>
#include <vector>

#include "fft_mpi_3d_api.h"          // the reference's (found through -I$(REF)/3dmpifft_opt/include)
#include "fast_transpose/kernels_120.h"
#include "fast_transpose/kernels_201.h"

extern "C" void hipcpu_set_device_count(int n);      // hipcpu.cpp

// ------------------------------------------------------------------------------------------ the FFT engine's entry points
typedef std::complex<double> cd;
static void dft_lines(cd* data, long n, long stride, long lines, long line_dist, bool inverse)
{
    std::vector<cd> w(n), x(n);
    for (long k = 0; k < n; k++) {
        const double a = (inverse ? 2.0 : -2.0) * M_PI * (double)k / (double)n;
        w[k] = cd(cos(a), sin(a));
    }
    for (long l = 0; l < lines; l++) {
        cd* p = data + l * line_dist;
        for (long j = 0; j < n; j++) x[j] = p[j * stride];
        for (long k = 0; k < n; k++) {
            cd acc(0, 0);
            for (long j = 0; j < n; j++) acc += x[j] * w[(k * j) % n];
            p[k * stride] = acc;
        }
    }
}

// Engine 1 (default when oracle/_ref/libtemplatefft_cpu.so is there): the reference's own generator + generated kernels
// (tfft_engine.cpp).  Engine 0: the plain DFT above.  The handle of engine 1 rides in app->localFFTPlan.
static int g_engine = -1;
static void* (*p_create)(int, long long, long long, long long, int) = nullptr;
static int (*p_launch)(void*, void**, int) = nullptr;
static void (*p_destroy)(void*) = nullptr;
static bool engine_available()
{
    static int state = -1;
    if (state < 0) {
        state = 0;
        Dl_info info;
        if (dladdr((void*)&engine_available, &info) && info.dli_fname) {
            std::string p(info.dli_fname);
            const size_t k = p.rfind('/');
            p = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/libtemplatefft_cpu.so";
            if (void* h = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL)) {
                p_create = (decltype(p_create))dlsym(h, "tfft_create");
                p_launch = (decltype(p_launch))dlsym(h, "tfft_launch");
                p_destroy = (decltype(p_destroy))dlsym(h, "tfft_destroy");
                state = p_create && p_launch && p_destroy;
            }
        }
    }
    return state == 1;
}
extern "C" int ref3d_set_engine(int e)        // 0 DFT, 1 the reference's engine; returns the engine in effect
{
    g_engine = (e == 1 && engine_available()) ? 1 : 0;
    return g_engine;
}
static int engine() { return g_engine < 0 ? ref3d_set_engine(1) : g_engine; }

FFTResult initializeFFT(FFTApplication* app, FFTConfiguration cfg)
{
    if (cfg.FFTdim < 1 || cfg.FFTdim > 2 || !cfg.doublePrecision) return FFT_ERROR_FAILED_TO_INITIALIZE;
    app->configuration = cfg;     // (cfg.bufferSize points at a local of setFFTPlans, api.cpp:396/425: never dereferenced here)
    app->localFFTPlan = nullptr;
    if (engine() == 1) {
        // null: a length the reference's generator cannot do (a prime factor > 7, tfft_engine.cpp); the DFT then stands in for this
        // application so that the surrounding reference code (tables, pack maps, exchange) can still be exercised on such sizes
        app->localFFTPlan = (FFTPlan*)p_create((int)cfg.FFTdim, (long long)cfg.size[0], (long long)cfg.size[1], (long long)cfg.size[2], cfg.makeInversePlanOnly ? 1 : 0);
    }
    return FFT_SUCCESS;
}
FFTResult setFFTArgs(GPU*, FFTApplication*, FFTLaunchParams*, int) { return FFT_SUCCESS; }
void deleteFFT(FFTApplication* app)
{
    if (app->localFFTPlan && p_destroy) p_destroy(app->localFFTPlan);
    app->localFFTPlan = nullptr;
}
hipError_t launchFFTKernel(FFTApplication* app, int inverse)
{
    const FFTConfiguration& c = app->configuration;
    if (app->localFFTPlan) return (hipError_t)p_launch(app->localFFTPlan, c.buffer, inverse);
    cd* data = (cd*)*c.buffer;
    const long s0 = (long)c.size[0], s1 = (long)(c.size[1] ? c.size[1] : 1), s2 = (long)(c.size[2] ? c.size[2] : 1);
    // axis 0: every line of the other two axes
    dft_lines(data, s0, 1, s1 * s2, s0, inverse != 0);
    if (c.FFTdim == 2)
        for (long b = 0; b < s2; b++) dft_lines(data + b * s0 * s1, s1, s0, s0, 1, inverse != 0);
    return hipSuccess;
}

// the reference's FFT engine on its own: an in-place transform of `data` (s0 fastest) over the first fftdim axes, the other axes
// being batches -- the templateFFT batch-test surface (templateFFT/batchTest/Test_1D.cpp, Test_2D.cpp).  -2: engine absent,
// -3: the generator refuses the size
extern "C" int ref3d_engine_fft(int fftdim, long long s0, long long s1, long long s2, int inverse, double* data)
{
    if (!engine_available()) return -2;
    void* h = p_create(fftdim, s0, s1, s2, inverse);
    if (!h) return -3;
    void* buf = data;
    const int rc = p_launch(h, &buf, inverse);
    p_destroy(h);
    return rc;
}

extern "C" int ref3d_engine_schedule(long long n, int* radices, int max_radices, int* uploads)
{
    if (!engine_available()) return -2;
    static int (*p_sched)(long long, int*, int, int*) = nullptr;
    if (!p_sched) {
        Dl_info info;
        dladdr((void*)p_create, &info);
        if (void* h = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD)) p_sched = (decltype(p_sched))dlsym(h, "tfft_schedule");
    }
    return p_sched ? p_sched(n, radices, max_radices, uploads) : -2;
}

// ------------------------------------------------------------------------------------------ cuTranspose dispatcher (restated)
extern "C" int cut_transpose3d(data_t* output, const data_t* input, const int* size, const int* permutation, int elements_per_thread)
{
    if (output == input || elements_per_thread != 1) return -1;     // fftX only uses the out-of-place, 1-element-per-thread form
    // (valid_parameters, transpose3d.cpp:352-377, also refuses any dimension < 2, which makes the reference abort when a device owns
    // a single y row; that restriction is a quirk the product and the oracle do not reproduce -- SURVEY A.4 -- and is left out)
    const int d2 = permutation[0] == 0 ? 1 : permutation[0];
    dim3 block(TILE_SIZE, TILE_SIZE / elements_per_thread, 1), grid;
    grid.x = (size[0] + TILE_SIZE - 1) / TILE_SIZE;
    grid.y = (size[d2] + TILE_SIZE - 1) / TILE_SIZE;
    grid.z = size[d2 == 1 ? 2 : 1];
    if (permutation[0] == 1 && permutation[1] == 2 && permutation[2] == 0)
        hipLaunchKernelGGL(dev_transpose_120_ept1, grid, block, 0, 0, output, input, size[0], size[1], size[2]);
    else if (permutation[0] == 2 && permutation[1] == 0 && permutation[2] == 1)
        hipLaunchKernelGGL(dev_transpose_201_ept1, grid, block, 0, 0, output, input, size[0], size[1], size[2]);
    else
        return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------ the driver's call sequence
// in[p], out[p]: getMaxDataCount elements each (interleaved doubles).  dumps (optional): dumps[(p * 4 + stage) * 2 + which]
// receives bufferDev1 (which 0) / bufferDev2 (which 1) of device p after stage 0..3 in execution order (forward: fftZY,
// localTransposeUneven, slabAlltoall, fftX; backward: fftX, slabAlltoall, localTransposeUneven, fftZY) -- then the four stage
// functions are called one by one with the barriers of fft_mpi_execute_dft_3d_c2c (api.cpp:181-214) around the exchange;
// without dumps the reference's own fft_mpi_execute_dft_3d_c2c runs.  Returns 0, or a negative code.
extern "C" int ref3d_run(int n0, int n1, int n2, int P, int direction, const double* const* in, double* const* out, double* const* dumps,
                         long long* tables /* P * P * 4: scount, soffset, rcount, roffset of every device, or null */)
{
    if (P < 1 || P > 64 || (direction != FORWARD && direction != BACKWARD)) return -1;
    hipcpu_set_device_count(P);
    const longInt64 N[3] = {n0, n1, n2};
    int newCount = 0, newCountInNode = 0;
    std::vector<longInt64> dataCount(P);
    fft_mpi_init(N, P, MPI_COMM_WORLD, newCount, newCountInNode, dataCount.data());     // api.cpp:3-39 (prints like the reference)
    if (newCount != P || newCountInNode != P) return -2;                                // the reference would run on fewer devices
    std::vector<Complex*> node_data(P, nullptr);
    int bad = 0;
    omp_set_dynamic(0);
#pragma omp parallel num_threads(P)
    {
        const int i = omp_get_thread_num();
        if (omp_get_num_threads() != P) {
#pragma omp atomic write
            bad = 1;
        } else {
            ROCM_CHECK(hipSetDevice(i));
            const bool last = i == P - 1;
            const longInt64 maxc = getMaxDataCount(n0, n1, n2, P, last);
            Complex* inDev = fft_mpi_alloc_local_memory((int)maxc, ALLOC_DEV);
            Complex* outDev = fft_mpi_alloc_local_memory((int)maxc, ALLOC_DEV);
            memcpy(inDev, in[i], (size_t)maxc * sizeof(Complex));
            fft_mpi_3d_plan_p plan = fft_mpi_plan_dft_c2c_3d(n0, n1, n2, inDev, outDev, node_data.data(), MPI_COMM_WORLD, i, P, P, direction);
            if (tables)
                for (int q = 0; q < P; q++) {
                    long long* t = tables + ((size_t)i * P + q) * 4;
                    t[0] = plan->tInfo.scount[q]; t[1] = plan->tInfo.soffset[q]; t[2] = plan->tInfo.rcount[q]; t[3] = plan->tInfo.roffset[q];
                }
#pragma omp barrier
            if (!dumps) {
                fft_mpi_execute_dft_3d_c2c(plan);
            } else {
                auto dump = [&](int stage) {
                    memcpy(dumps[((size_t)i * 4 + stage) * 2 + 0], plan->bufferDev1, (size_t)maxc * sizeof(Complex));
                    memcpy(dumps[((size_t)i * 4 + stage) * 2 + 1], plan->bufferDev2, (size_t)maxc * sizeof(Complex));
                };
                if (direction == FORWARD) {
                    fftZY(plan); dump(0);
                    localTransposeUneven(plan); dump(1);
#pragma omp barrier
                    slabAlltoall(plan);
#pragma omp barrier
                    dump(2);
                    fftX(plan); dump(3);
                } else {
                    fftX(plan); dump(0);
#pragma omp barrier
                    slabAlltoall(plan);
#pragma omp barrier
                    dump(1);
                    localTransposeUneven(plan); dump(2);
                    fftZY(plan); dump(3);
                }
            }
#pragma omp barrier
            memcpy(out[i], outDev, (size_t)maxc * sizeof(Complex));     // bufferDev2 == outDev (api.cpp:68-75)
            deleteFFT(&plan->appYZ);          // (fft_mpi_destroy_plan, api.cpp:143-179, leaves the two FFT applications behind)
            deleteFFT(&plan->appX);
            fft_mpi_destroy_plan(plan);
            hipFree(inDev);
            hipFree(outDev);
        }
    }
    return bad ? -3 : 0;
}

// plan creation only (api.cpp:41-141): the TransInfo tables of every device, for geometries too large to transform with a DFT
extern "C" int ref3d_tables(int n0, int n1, int n2, int P, int direction, long long* tables)
{
    if (P < 1 || P > 64) return -1;
    hipcpu_set_device_count(P);
    std::vector<Complex*> node_data(P, nullptr);
    for (int i = 0; i < P; i++) {
        const longInt64 maxc = getMaxDataCount(n0, n1, n2, P, i == P - 1);
        Complex* inDev = fft_mpi_alloc_local_memory((int)maxc, ALLOC_DEV);
        fft_mpi_3d_plan_p plan = fft_mpi_plan_dft_c2c_3d(n0, n1, n2, inDev, nullptr, node_data.data(), MPI_COMM_WORLD, i, P, P, direction);
        for (int q = 0; q < P; q++) {
            long long* t = tables + ((size_t)i * P + q) * 4;
            t[0] = plan->tInfo.scount[q]; t[1] = plan->tInfo.soffset[q]; t[2] = plan->tInfo.rcount[q]; t[3] = plan->tInfo.roffset[q];
        }
        deleteFFT(&plan->appYZ);
        deleteFFT(&plan->appX);
        fft_mpi_destroy_plan(plan);
        hipFree(inDev);
    }
    return 0;
}

// the count helpers, straight from the reference (api.cpp:289-316, 232-287)
extern "C" long long ref3d_max_data_count(int n0, int n1, int n2, int P, int is_last) { return getMaxDataCount(n0, n1, n2, P, is_last != 0); }
extern "C" int ref3d_proper_device_num(long long n0, int wanted, int have)
{
    hipcpu_set_device_count(have);
    const longInt64 N[3] = {n0, 1, 1};
    int total = 0, in_node = 0;
    getProperDeviceNum(N, wanted, 1, 0, total, in_node);
    return total;
}
