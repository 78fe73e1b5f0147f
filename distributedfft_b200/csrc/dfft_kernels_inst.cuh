// dfft_kernels_inst.cuh -- launcher template + size-table macros, included by the per-precision
// instantiation units.
#pragma once
#include <atomic>
#include "dfft_kernels.cuh"

namespace dfft {

template <class S, typename T, int C, int MI, int MO, bool TW, bool CI, bool CO, int MINB, bool PP>
cudaError_t launch_pass(const void* vargs, int sm_count, cudaStream_t st)
{
    const TileArgs<T>& a = *reinterpret_cast<const TileArgs<T>*>(vargs);
    auto kern = fft_tile_kernel<S, T, C, MI, MO, TW, CI, CO, MINB, PP>;
    constexpr size_t smem = TileSmem<S, T, C, PP>::bytes(CI || CO);
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int occ = occ_cache[dev & 63].load();
    if (occ == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, S::T * C, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        occ_cache[dev & 63].store(occ);
    }
    if (a.ntiles <= 0) return cudaSuccess;
    long long grid = (long long)sm_count * occ;
    if (grid > a.ntiles) grid = a.ntiles;
    kern<<<(unsigned)grid, S::T * C, smem, st>>>(a);
    return cudaGetLastError();
}

template <class S> void fill_rad(int& n, int* rad)
{
    n = S::NSTAGES;
    for (int i = 0; i < S::NSTAGES; i++) rad[i] = S::rad(i);
}

// ZS/SS: schedules; ZC/SC lines per tile; *TW twiddles in registers; *MB min CTAs/SM; *PP ping-pong smem
template <typename T, class ZS, int ZC, bool ZTW, int ZMB, bool ZPP, class SS, int SC, bool STW, int SMB, bool SPP>
SizeEntry make_entry()
{
    SizeEntry e{};
    e.N = ZS::N;
    e.prec = sizeof(T) == 8 ? 0 : 1;
    e.z_C = ZC;
    e.s_C = SC;
    fill_rad<ZS>(e.z_nstages, e.z_rad);
    fill_rad<SS>(e.s_nstages, e.s_rad);
    e.launch[PK_Z] = launch_pass<ZS, T, ZC, MAP_T, MAP_T, ZTW, false, false, ZMB, ZPP>;
    e.launch[PK_Y] = launch_pass<SS, T, SC, MAP_C, MAP_C, STW, false, false, SMB, SPP>;
    e.launch[PK_Y_CO] = launch_pass<SS, T, SC, MAP_C, MAP_C, STW, false, true, SMB, SPP>;
    e.launch[PK_Y_CI] = launch_pass<SS, T, SC, MAP_C, MAP_C, STW, true, false, SMB, SPP>;
    e.launch[PK_XF] = launch_pass<SS, T, SC, MAP_C, MAP_T, STW, false, false, SMB, SPP>;
    e.launch[PK_XB] = launch_pass<SS, T, SC, MAP_T, MAP_C, STW, false, false, SMB, SPP>;
    e.launch[PK_XB_CO] = launch_pass<SS, T, SC, MAP_T, MAP_C, STW, false, true, SMB, SPP>;
    return e;
}

}  // namespace dfft
