// distFFT -- the speedTest driver (3dmpifft_opt/fftSpeed3d_c2c.cpp:8-143 re-done for CUDA, one host
// thread per GPU instead of `#pragma omp parallel for`, no MPI):
//     distFFT NX NY NZ GPU_COUNT [--json]
// Same call sequence, same ramp input, same round-trip error metric (/1e7), same report block; adds
// the absolute round-trip error, per-stage milliseconds and the HBM-roofline fraction.
#include <cuda_runtime.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "fft_mpi_3d_api.h"

#define CUDA_CHECK(stmt)                                                                                          \
    do {                                                                                                          \
        cudaError_t e_ = (stmt);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            fprintf(stderr, "[%s:%d] CUDA call '%s' failed with %d: %s \n", __FILE__, __LINE__, #stmt, e_, cudaGetErrorString(e_)); \
            exit(EXIT_FAILURE);                                                                                   \
        }                                                                                                         \
    } while (0)

int main(int argc, char* argv[])
{
    char hostname[256];
    gethostname(hostname, sizeof(hostname));
    printf("PID %d on %s ready for attach\n", getpid(), hostname);
    fflush(stdout);
    bool json = false;
    if (argc == 6 && !strcmp(argv[5], "--json")) { json = true; argc = 5; }
    if (argc != 5) {
        printf("The format of arguments should be [NX, NY, NZ, GPU_COUNT]!\n");
        exit(EXIT_FAILURE);
    }
    int devCount;
    CUDA_CHECK(cudaGetDeviceCount(&devCount));
    const longInt64 N[3] = {atoll(argv[1]), atoll(argv[2]), atoll(argv[3])};
    const int iniDeviceNumInNode = atoi(argv[4]);
    if (iniDeviceNumInNode < 1) { printf("GPU_COUNT must be >= 1\n"); exit(EXIT_FAILURE); }
    int newDeviceCount, newDeviceCountInNode;
    std::vector<longInt64> dataCountInNode(iniDeviceNumInNode);
    fft_mpi_init(N, iniDeviceNumInNode, MPI_COMM_WORLD, newDeviceCount, newDeviceCountInNode, dataCountInNode.data());
    const int deviceCountInNode = newDeviceCountInNode, totalDeviceCount = newDeviceCount;

    std::vector<Complex*> node_data_dev(deviceCountInNode, nullptr);
    double maxErrInProcess = 1e-30, maxAbsErr = 0, forwardTimeProcess = 1e-30, eventTimeProcess = 0;
    double stage[5] = {0, 0, 0, 0, 0};
    std::mutex mu;
    auto worker = [&](int i) {
        int globalIdx = i;
        CUDA_CHECK(cudaSetDevice(globalIdx % devCount));
        const longInt64 normalDeviceDataCount = (longInt64)std::ceil((double)N[0] / totalDeviceCount) * N[1] * N[2];
        const longInt64 cnt = dataCountInNode[i];
        Complex* data_cpu = (Complex*)malloc(cnt * sizeof(Complex));
        Complex* data_cpu_out = (Complex*)malloc(cnt * sizeof(Complex));
        for (longInt64 j = 0; j < cnt; ++j) data_cpu[j][0] = data_cpu[j][1] = (double)(i * normalDeviceDataCount + j);   // drv.cpp:61-63
        bool isLastDev = globalIdx == totalDeviceCount - 1;
        longInt64 maxDataCountDev = getMaxDataCount((int)N[0], (int)N[1], (int)N[2], totalDeviceCount, isLastDev);
        Complex* inDev = fft_mpi_alloc_local_memory(maxDataCountDev, ALLOC_DEV);
        Complex* outDev = fft_mpi_alloc_local_memory(maxDataCountDev, ALLOC_DEV);
        CUDA_CHECK(cudaMemcpy(inDev, data_cpu, cnt * sizeof(Complex), cudaMemcpyHostToDevice));
        fft_mpi_3d_plan_p plan = fft_mpi_plan_dft_c2c_3d(N[0], N[1], N[2], inDev, outDev, node_data_dev.data(), MPI_COMM_WORLD, i,
                                                         deviceCountInNode, totalDeviceCount, FORWARD);
        CUDA_CHECK(cudaMemcpy(plan->bufferDev1, data_cpu, cnt * sizeof(Complex), cudaMemcpyHostToDevice));   // drv.cpp:78
        fft_mpi_execute_dft_3d_c2c(plan);
        fft_mpi_3d_plan_p planBack = fft_mpi_plan_dft_c2c_3d(N[0], N[1], N[2], outDev, inDev, node_data_dev.data(), MPI_COMM_WORLD, i,
                                                             deviceCountInNode, totalDeviceCount, BACKWARD);
        fft_mpi_execute_dft_3d_c2c(planBack);
        CUDA_CHECK(cudaMemcpy(data_cpu_out, inDev, cnt * sizeof(Complex), cudaMemcpyDeviceToHost));
        double maxErr = -1.0;
        const double n3 = (double)N[0] * (double)N[1] * (double)N[2];
        for (longInt64 j = 0; j < cnt; ++j) {   // drv.cpp:84-91
            double tmp1 = data_cpu[j][0] - data_cpu_out[j][0] / n3, tmp2 = data_cpu[j][1] - data_cpu_out[j][1] / n3;
            double err = std::sqrt(tmp1 * tmp1 + tmp2 * tmp2);
            if (maxErr < err) maxErr = err;
        }
        fft_mpi_execute_dft_3d_c2c(plan);
        auto t0 = std::chrono::steady_clock::now();
        fft_mpi_execute_dft_3d_c2c(plan);   // synchronous (waits for the device) like the reference
        double forward_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // the report block prints the host wall clock of the synchronous call like the reference (drv.cpp:94-98);
        // the device-side time of the same call (CUDA events on the plan's stream) is an extra line
        double ev_total = plan->t[4] * 1e-3;
        double st[5];
        memcpy(st, plan->t, sizeof(st));
        fft_mpi_execute_dft_3d_c2c(plan);
        fft_mpi_destroy_plan(plan);
        fft_mpi_destroy_plan(planBack);
        {
            std::lock_guard<std::mutex> lk(mu);
            if (maxErrInProcess < maxErr / 1e7) maxErrInProcess = maxErr / 1e7;
            if (maxAbsErr < maxErr) maxAbsErr = maxErr;
            if (forwardTimeProcess < forward_time) forwardTimeProcess = forward_time;
            if (eventTimeProcess < ev_total) eventTimeProcess = ev_total;
            for (int k = 0; k < 5; k++) if (stage[k] < st[k]) stage[k] = st[k];
        }
        free(data_cpu_out);
        free(data_cpu);
        CUDA_CHECK(cudaFree(inDev));
        CUDA_CHECK(cudaFree(outDev));
    };
    std::vector<std::thread> th;
    for (int i = 0; i < deviceCountInNode; ++i) th.emplace_back(worker, i);
    for (auto& t : th) t.join();

    long long fftsize = N[0] * N[1] * N[2];
    double gflops = 5.0 * fftsize * std::log((double)fftsize) * 1e-9 / std::log(2.0) / forwardTimeProcess;
    std::cout << "\n----------------------------------------------------------------------------- \n";
    std::cout << "distributed FFT performance test\n";
    std::cout << "----------------------------------------------------------------------------- \n";
    std::cout << "Size:             " << N[0] << "x" << N[1] << "x" << N[2] << "\n";
    std::cout << "MPI ranks:        " << totalDeviceCount << "\n";
    std::cout << "Forward FFT time: " << forwardTimeProcess << " (s)\n";
    std::cout << "Performance:      " << gflops << " GFlops/s\n";
    std::cout << "Max error:        " << maxErrInProcess << "\n";
    std::cout << std::endl;
    // additions (not in the reference report)
    const double M = (double)fftsize / totalDeviceCount;
    const double bytes = (6.0 + (totalDeviceCount > 1 ? 2.0 : 0.0)) * 16.0 * M;   // (6 + 2*[P>1]) * E * M
    std::cout << "Max abs error:    " << maxAbsErr << " (round trip, unscaled by 1e7)\n";
    std::cout << "Stage ms:         t0 " << stage[0] << "  t1 " << stage[1] << "  t2 " << stage[2] << "  t3 " << stage[3] << "\n";
    const double devTime = eventTimeProcess > 0 ? eventTimeProcess : forwardTimeProcess;
    std::cout << "Device time:      " << devTime << " (s) (CUDA events, max over devices) -> "
              << 5.0 * fftsize * std::log2((double)fftsize) * 1e-9 / devTime << " GFlops/s\n";
    std::cout << "HBM traffic:      " << bytes * 1e-9 << " GB algorithmic per GPU -> " << bytes / devTime * 1e-9 << " GB/s per GPU\n";
    if (json)
        printf("{\"size\": [%lld, %lld, %lld], \"gpus\": %d, \"forward_s\": %.9g, \"device_s\": %.9g, \"gflops\": %.6g, \"max_error\": %.6g, \"max_abs_error\": %.6g, "
               "\"t0_ms\": %.6g, \"t1_ms\": %.6g, \"t2_ms\": %.6g, \"t3_ms\": %.6g}\n",
               N[0], N[1], N[2], totalDeviceCount, forwardTimeProcess, devTime, gflops, maxErrInProcess, maxAbsErr, stage[0], stage[1], stage[2], stage[3]);
    return 0;
}
