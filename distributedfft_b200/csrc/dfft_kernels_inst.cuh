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

template <class OpA, class OpB, typename T, int MINB>
cudaError_t launch_fused(const void* va, const void* vb, const FusedCtl* ctl, int sm_count, cudaStream_t st)
{
    const TileArgs<T>& a = *reinterpret_cast<const TileArgs<T>*>(va);
    const TileArgs<T>& b = *reinterpret_cast<const TileArgs<T>*>(vb);
    auto kern = fft_fused2_kernel<OpA, OpB, T, MINB>;
    constexpr size_t exch = OpA::SM::exch_bytes > OpB::SM::exch_bytes ? OpA::SM::exch_bytes : OpB::SM::exch_bytes;
    constexpr size_t smem = exch + OpA::aux_bytes + OpB::aux_bytes;
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int occ = occ_cache[dev & 63].load();
    if (occ == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OpA::NT, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        occ_cache[dev & 63].store(occ);
    }
    if (ctl->planes <= 0) return cudaSuccess;
    const long long total = ctl->planes * ((long long)ctl->GA + ctl->GB);
    long long grid = (long long)sm_count * occ;
    if (grid > total) grid = total;
    FusedCtl c = *ctl;
    if (c.lag <= 0) c.lag = (int)((grid + c.GA + c.GB - 1) / ((long long)c.GA + c.GB)) + 1;   // role A stays one in-flight window ahead
    kern<<<(unsigned)grid, OpA::NT, smem, st>>>(a, b, c);
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
    // fused two-pass kernels: the contiguous role is re-tiled so that both roles fill the same CTA
    constexpr int NT = SS::T * SC;
    static_assert(NT % ZS::T == 0, "strided CTA size must be a multiple of the contiguous line's thread count");
    constexpr int FZC = NT / ZS::T;
    e.f_zC = FZC;
    using OZ = TileOp<ZS, T, FZC, MAP_T, MAP_T, false, false, false, false>;
    using OY = TileOp<SS, T, SC, MAP_C, MAP_C, false, false, false, false>;
    using OYco = TileOp<SS, T, SC, MAP_C, MAP_C, false, false, true, false>;
    using OYci = TileOp<SS, T, SC, MAP_C, MAP_C, false, true, false, false>;
    e.fused[FK_ZY] = launch_fused<OZ, OY, T, SMB>;
    e.fused[FK_ZY_CO] = launch_fused<OZ, OYco, T, SMB>;
    e.fused[FK_YZ] = launch_fused<OY, OZ, T, SMB>;
    e.fused[FK_YZ_CI] = launch_fused<OYci, OZ, T, SMB>;
    return e;
}

}  // namespace dfft
