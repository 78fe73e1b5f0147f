// tma_t0.cu -- developer prototype (not part of the product path): TMA-pipelined pass kernels and a TMA-pipelined fused
// t0 (Z then Y through L2) for 512-point fp64 lines.  One CTA per SM; a 3-slot shared-memory ring of 64 KB tiles
// (8 lines x 512 points): a tile is fetched by TMA (1-D bulk copy for contiguous lines, 2-D tensor copies for columns),
// transformed IN PLACE in its slot by 512 consumer threads (exchanges through the slot itself, XOR-swizzled for the
// line-major map), and written back by a TMA store, so global-memory latency never sits on the compute threads.
//   tma_t0 [planes]        runs Z pass, Y pass, fused t0; checks each on plane-wave inputs; prints ms and GB/s
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fft_core.cuh"
#include "../fft_tma.cuh"

using namespace dfft;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

using S = Sched<512, 8, 8, 8, 8>;
constexpr int N = 512, C = 8, TT = S::T, NCONS = TT * C;   // 64 threads per line, 512 consumer threads
constexpr int TILE_ELEMS = N * C;
constexpr uint32_t TILE_BYTES = TILE_ELEMS * 16;
constexpr int NSLOT = 3;
constexpr int LUT_ENTRIES = S::lut_size();
constexpr size_t SMEM_BYTES = (size_t)NSLOT * TILE_BYTES + (size_t)((LUT_ENTRIES * 16 + 127) / 128 * 128) + 256;

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tensor_g2s_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor_g2s_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor_s2g_2d(const CUtensorMap* map, int c0, int c1, const void* src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int NLEFT> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NLEFT) : "memory"); }
template <int NLEFT> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(NLEFT) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// ---- in-place transform of one tile in its slot ---------------------------------------------------------------------
// LINE = true : slot holds C contiguous lines   [c][pos]      (Z role), exchanges XOR-swizzled inside each line
// LINE = false: slot holds C columns            [pos][c]      (Y role), conflict-free as is
template <bool LINE> __device__ __forceinline__ int nat(int c, int pos) { return LINE ? c * N + pos : pos * C + c; }
template <bool LINE> __device__ __forceinline__ int swz(int c, int pos) { return LINE ? c * N + (pos ^ ((pos >> 3) & 7)) : pos * C + c; }

template <bool LINE, int s> __device__ __forceinline__ void exchange(double2* v, double2* buf, int t, int c)
{
    constexpr int RAD = S::rad(s), NS = S::ns(s);
    static_assert(S::R / RAD == 1, "one butterfly per thread");
    const int k = t % NS, j0 = (t - k) * RAD + k;
#pragma unroll
    for (int m = 0; m < RAD; m++) buf[swz<LINE>(c, j0 + m * NS)] = v[m];
    cons_sync();
#pragma unroll
    for (int u = 0; u < S::R; u++) v[u] = buf[swz<LINE>(c, t + u * TT)];
}

template <bool LINE> __device__ __forceinline__ void transform_tile(double2* buf, const double2* lut, int tid)
{
    int t, c;
    if (LINE) { t = tid % TT; c = tid / TT; } else { c = tid % C; t = tid / C; }
    double2 v[S::R];
#pragma unroll
    for (int u = 0; u < S::R; u++) v[u] = buf[nat<LINE>(c, t + u * TT)];
    stage_compute<S, 0, double, false>(v, t, lut, nullptr);
    cons_sync();                       // everyone has read its inputs: the slot is free for the exchanges
    exchange<LINE, 0>(v, buf, t, c);
    stage_compute<S, 1, double, false>(v, t, lut, nullptr);
    cons_sync();                       // all gathers of the previous exchange are done
    exchange<LINE, 1>(v, buf, t, c);
    stage_compute<S, 2, double, false>(v, t, lut, nullptr);
    cons_sync();
#pragma unroll
    for (int u = 0; u < S::R; u++) buf[nat<LINE>(c, t + u * TT)] = v[u];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes visible to the TMA store
    cons_sync();
}

// X role: the tile arrives as C columns [pos][c] and leaves as C contiguous lines [c][pos]: the thread map changes from
// column-major to line-major in the last exchange, whose buffer is XOR-swizzled so that both sides are conflict free
__device__ __forceinline__ void transform_tile_x(double2* buf, const double2* lut, int tid)
{
    const int c_in = tid % C, t_in = tid / C, t_out = tid % TT, c_out = tid / TT;
    double2 v[S::R];
#pragma unroll
    for (int u = 0; u < S::R; u++) v[u] = buf[(t_in + u * TT) * C + c_in];
    stage_compute<S, 0, double, false>(v, t_in, lut, nullptr);
    cons_sync();
    exchange<false, 0>(v, buf, t_in, c_in);
    stage_compute<S, 1, double, false>(v, t_in, lut, nullptr);
    cons_sync();
    {
        constexpr int RAD = S::rad(1), NS = S::ns(1);
        const int k = t_in % NS, j0 = (t_in - k) * RAD + k;
#pragma unroll
        for (int m = 0; m < RAD; m++) { const int pos = j0 + m * NS; buf[pos * C + (c_in ^ (pos & 7))] = v[m]; }
        cons_sync();
#pragma unroll
        for (int u = 0; u < S::R; u++) { const int pos = t_out + u * TT; v[u] = buf[pos * C + (c_out ^ (pos & 7))]; }
    }
    stage_compute<S, 2, double, false>(v, t_out, lut, nullptr);
    cons_sync();
#pragma unroll
    for (int u = 0; u < S::R; u++) buf[c_out * N + t_out + u * TT] = v[u];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cons_sync();
}

struct Smem {
    unsigned char* base;
    __device__ __forceinline__ double2* slot(int i) const { return reinterpret_cast<double2*>(base + (size_t)i * TILE_BYTES); }
    double2* lut;
    uint64_t* full;    // [NSLOT]
    uint64_t* empty;   // [NSLOT]
    int* desc;         // [NSLOT][4]
};
__device__ __forceinline__ Smem carve(unsigned char* raw)
{
    Smem s;
    s.base = raw;
    unsigned char* p = raw + (size_t)NSLOT * TILE_BYTES;
    s.lut = reinterpret_cast<double2*>(p);
    p += (LUT_ENTRIES * 16 + 127) / 128 * 128;
    s.full = reinterpret_cast<uint64_t*>(p);
    s.empty = s.full + NSLOT;
    s.desc = reinterpret_cast<int*>(s.empty + NSLOT);
    return s;
}

// ---- pass kernels (static tile assignment): MODE 0 = contiguous lines (Z), 1 = columns (Y) --------------------------------
// Z: tile i = lines [8 i, 8 i + 8) of a dense array;  Y: tile i = (plane a = i / 64, column group b = i % 64)
template <int MODE>
__global__ void __launch_bounds__(NCONS, 1) pass_tma(const double2* in, double2* out, const __grid_constant__ CUtensorMap map_in,
                                                      const __grid_constant__ CUtensorMap map_out, const double2* lut_g, long long ntiles)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Smem sm = carve(raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSLOT; i++) mbar_init(sm.full + i, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < LUT_ENTRIES; i += NCONS) sm.lut[i] = lut_g[i];
    __syncthreads();
    const long long mine = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_of = [&](long long i) { return (long long)blockIdx.x + i * gridDim.x; };
    auto load = [&](long long i) {
        const int s = (int)(i % NSLOT);
        const long long tile = tile_of(i);
        mbar_expect_tx(sm.full + s, TILE_BYTES);
        if (MODE == 0) bulk_g2s(sm.slot(s), in + tile * TILE_ELEMS, TILE_BYTES, sm.full + s);
        else if (MODE == 1) {
            const int a = (int)(tile / (N / C)), b = (int)(tile % (N / C));
            tensor_g2s_2d(sm.slot(s), &map_in, b * C * 2, a * N, sm.full + s);
            tensor_g2s_2d(sm.slot(s) + TILE_ELEMS / 2, &map_in, b * C * 2, a * N + N / 2, sm.full + s);
        } else {   // X: tile (y = a, z group b): rows x = 0..N-1 of the 3-D tensor (z, y, x)
            const int a = (int)(tile / (N / C)), b = (int)(tile % (N / C));
            tensor_g2s_3d(sm.slot(s), &map_in, b * C * 2, a, 0, sm.full + s);
            tensor_g2s_3d(sm.slot(s) + TILE_ELEMS / 2, &map_in, b * C * 2, a, N / 2, sm.full + s);
        }
    };
    if (tid == 0)
        for (long long i = 0; i < mine && i < NSLOT - 1; i++) load(i);
    for (long long i = 0; i < mine; i++) {
        const int s = (int)(i % NSLOT);
        mbar_wait(sm.full + s, (uint32_t)((i / NSLOT) & 1));
        if (MODE == 0) transform_tile<true>(sm.slot(s), sm.lut, tid);
        else if (MODE == 1) transform_tile<false>(sm.slot(s), sm.lut, tid);
        else transform_tile_x(sm.slot(s), sm.lut, tid);
        if (tid == 0) {
            const long long tile = tile_of(i);
            if (MODE == 0 || MODE == 2) bulk_s2g(out + tile * TILE_ELEMS, sm.slot(s), TILE_BYTES);   // X: out[(y N + z) N + x], tile = y * 64 + z / 8
            else {
                const int a = (int)(tile / (N / C)), b = (int)(tile % (N / C));
                tensor_s2g_2d(&map_out, b * C * 2, a * N, sm.slot(s));
                tensor_s2g_2d(&map_out, b * C * 2, a * N + N / 2, sm.slot(s) + TILE_ELEMS / 2);
            }
            bulk_commit();
            if (i + NSLOT - 1 < mine) {
                bulk_wait_read<1>();          // the store of tile i-1 has drained its slot, which tile i+2 reuses
                load(i + NSLOT - 1);
            }
        }
    }
    if (tid == 0) bulk_wait<0>();
}

// ---- fused t0: Z role (src -> mid) and Y role (mid -> mid, in place), tickets + per-plane completion counters -------------
struct FusedT0 {
    const double2* src;
    double2* mid;
    unsigned long long* plane_done;
    unsigned int* ticket;
    unsigned long long target;     // plane_done value meaning "all 64 Z tiles of the plane are stored" in this launch
    int planes, lag;
};

__global__ void __launch_bounds__(NCONS + 32, 1) fused_t0_tma(const FusedT0 F, const __grid_constant__ CUtensorMap map_mid, const double2* lut_g)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Smem sm = carve(raw);
    const int tid = threadIdx.x;
    constexpr int G = N / C;   // 64 tiles per plane and role
    if (tid == 0) {
        for (int i = 0; i < NSLOT; i++) { mbar_init(sm.full + i, 1); mbar_init(sm.empty + i, 1); }
        fence_barrier_init();
    }
    for (int i = tid; i < LUT_ENTRIES; i += blockDim.x) sm.lut[i] = lut_g[i];
    __syncthreads();
    const long long lag = F.lag < F.planes ? F.lag : F.planes;
    const long long headT = lag * G, midT = (F.planes - lag) * (2 * G), total = headT + midT + lag * G;

    if (tid >= NCONS) {
        // ---------------- producer warp: one lane runs ahead of the consumers, up to NSLOT tiles --------------------------
        if (tid == NCONS) {
            for (long long n = 0;; n++) {
                const int s = (int)(n % NSLOT);
                if (n >= NSLOT) mbar_wait(sm.empty + s, (uint32_t)(((n / NSLOT) - 1) & 1));   // slot drained by its previous store
                const long long t = (long long)atomicAdd(F.ticket, 1u);
                int role, plane, idx;
                if (t >= total) { role = -1; plane = 0; idx = 0; }
                else if (t < headT) { role = 0; plane = (int)(t / G); idx = (int)(t % G); }
                else if (t < headT + midT) {
                    const long long u = t - headT, i = u / (2 * G), r = u % (2 * G);
                    if (r < G) { role = 0; plane = (int)(lag + i); idx = (int)r; }
                    else { role = 1; plane = (int)i; idx = (int)(r - G); }
                } else { const long long u = t - headT - midT; role = 1; plane = (int)(F.planes - lag + u / G); idx = (int)(u % G); }
                sm.desc[s * 4 + 0] = role; sm.desc[s * 4 + 1] = plane; sm.desc[s * 4 + 2] = idx;
                if (role < 0) { mbar_arrive(sm.full + s); break; }   // end marker (no bytes)
                if (role == 1) {                                    // the plane must have been written by the Z role
                    unsigned long long v;
                    unsigned spins = 0;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(F.plane_done + plane) : "memory");
                        if (v >= F.target) break;
                        __nanosleep(64);
                        if (++spins > (1u << 26)) __trap();
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");   // the acquired data is read through the async proxy
                }
                mbar_expect_tx(sm.full + s, TILE_BYTES);
                if (role == 0) bulk_g2s(sm.slot(s), F.src + ((long long)plane * G + idx) * TILE_ELEMS, TILE_BYTES, sm.full + s);
                else {
                    tensor_g2s_2d(sm.slot(s), &map_mid, idx * C * 2, plane * N, sm.full + s);
                    tensor_g2s_2d(sm.slot(s) + TILE_ELEMS / 2, &map_mid, idx * C * 2, plane * N + N / 2, sm.full + s);
                }
            }
        }
        return;
    }
    // ---------------- consumers ------------------------------------------------------------------------------------------
    int prev_role = -1, prev_plane = 0, prev_slot = 0;
    auto retire_prev = [&]() {   // thread 0: the previous tile's store is complete -> free its slot, publish a finished Z tile
        mbar_arrive(sm.empty + prev_slot);
        if (prev_role == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            atomicAdd(F.plane_done + prev_plane, 1ull);
        }
        prev_role = -1;
    };
    for (long long n = 0;; n++) {
        const int s = (int)(n % NSLOT);
        if (tid == 0) {
            // if the next tile has not landed yet, use the time to retire the previous one NOW: its load may be waiting on
            // exactly that Z tile (deferring the signal behind a dependent tile would deadlock)
            while (!mbar_try_wait(sm.full + s, (uint32_t)((n / NSLOT) & 1)))
                if (prev_role >= 0) { bulk_wait<0>(); retire_prev(); }
        }
        mbar_wait(sm.full + s, (uint32_t)((n / NSLOT) & 1));
        const int role = sm.desc[s * 4 + 0], plane = sm.desc[s * 4 + 1], idx = sm.desc[s * 4 + 2];
        if (role < 0) break;
        if (role == 0) transform_tile<true>(sm.slot(s), sm.lut, tid);
        else transform_tile<false>(sm.slot(s), sm.lut, tid);
        if (tid == 0) {
            if (role == 0) bulk_s2g(F.mid + ((long long)plane * G + idx) * TILE_ELEMS, sm.slot(s), TILE_BYTES);
            else {
                tensor_s2g_2d(&map_mid, idx * C * 2, plane * N, sm.slot(s));
                tensor_s2g_2d(&map_mid, idx * C * 2, plane * N + N / 2, sm.slot(s) + TILE_ELEMS / 2);
            }
            bulk_commit();
            if (prev_role >= 0) {
                bulk_wait<1>();                                   // the previous tile's store is complete (and has left its slot)
                retire_prev();
            }
            prev_role = role; prev_plane = plane; prev_slot = s;
        }
    }
    if (tid == 0) {
        bulk_wait<0>();
        if (prev_role >= 0) retire_prev();
        const unsigned left = atomicAdd(F.ticket + 1, 1u);
        if (left == gridDim.x - 1) { F.ticket[0] = 0; F.ticket[1] = 0; __threadfence(); }
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(void* base, long long rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q));
        if (!fn) { printf("cuTensorMapEncodeTiled not available\n"); exit(1); }
    }
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)N * 2, (cuuint64_t)rows};        // doubles per row, rows
    cuuint64_t strides[1] = {(cuuint64_t)N * 16};                      // bytes between rows
    cuuint32_t box[2] = {(cuuint32_t)C * 2, (cuuint32_t)N / 2};        // 8 complex columns x 256 rows = 32 KB
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

static CUtensorMap make_map3(void* base, int nx)
{
    EncodeTiledFn fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q));
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)N * 2, (cuuint64_t)N, (cuuint64_t)nx};           // z (doubles), y, x
    cuuint64_t strides[2] = {(cuuint64_t)N * 16, (cuuint64_t)N * N * 16};
    cuuint32_t box[3] = {(cuuint32_t)C * 2, 1, (cuuint32_t)N / 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled (3-D) failed: %d\n", (int)r); exit(1); }
    return m;
}

// X: element (x, y, z) = wave along x with frequency f = (y + z) % N; output layout [y][z][x]
__global__ void fill_x(double2* a)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)N * N * N) return;
    const int z = (int)(i % N), y = (int)((i / N) % N), x = (int)(i / ((long long)N * N));
    double sn, cs; sincospi(2.0 * (double)((long long)((y + z) % N) * x % N) / N, &sn, &cs);
    a[i] = make_double2(cs, sn);
}
__global__ void check_x(const double2* a, double* maxerr)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)N * N * N) return;
    const int x = (int)(i % N), z = (int)((i / N) % N), y = (int)(i / ((long long)N * N));
    const double ex = x == (y + z) % N ? (double)N : 0.0;
    const double err = fmax(fabs(a[i].x - ex), fabs(a[i].y));
    if (err > 1e-9) atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(err));
}

// plane wave per line / per plane; after the transform the energy sits in one bin
__global__ void fill_lines(double2* a, long long nlines)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nlines * N) return;
    const long long line = i / N; const int e = (int)(i % N);
    const int f = (int)(line % N);
    double sn, cs; sincospi(2.0 * (double)((long long)f * e % N) / N, &sn, &cs);
    a[i] = make_double2(cs, sn);
}
__global__ void check_lines(const double2* a, long long nlines, double* maxerr)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nlines * N) return;
    const long long line = i / N; const int e = (int)(i % N);
    const double ex = e == (int)(line % N) ? (double)N : 0.0;
    const double err = fmax(fabs(a[i].x - ex), fabs(a[i].y));
    if (err > 1e-9) atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(err));
}
// columns: element (plane, y, z) = wave along y with frequency f = (plane + z) % N
__global__ void fill_cols(double2* a, int planes)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)planes * N * N) return;
    const int z = (int)(i % N), y = (int)((i / N) % N), p = (int)(i / ((long long)N * N));
    const int f = (p + z) % N;
    double sn, cs; sincospi(2.0 * (double)((long long)f * y % N) / N, &sn, &cs);
    a[i] = make_double2(cs, sn);
}
__global__ void check_cols(const double2* a, int planes, double* maxerr)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)planes * N * N) return;
    const int z = (int)(i % N), y = (int)((i / N) % N), p = (int)(i / ((long long)N * N));
    const double ex = y == (p + z) % N ? (double)N : 0.0;
    const double err = fmax(fabs(a[i].x - ex), fabs(a[i].y));
    if (err > 1e-9) atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(err));
}
__global__ void fill_random(double2* a, long long n)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    a[i] = make_double2((double)(s & 0xffffff) * (1.0 / 16777216.0), (double)((s >> 24) & 0xffffff) * (1.0 / 16777216.0));
}
// 2-D: plane p holds e^{2 pi i (fy y + fz z)/N}, fy = 7p % N, fz = 13p % N  ->  N^2 at (fy, fz)
__global__ void fill_2d(double2* a, int planes)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)planes * N * N) return;
    const int z = (int)(i % N), y = (int)((i / N) % N), p = (int)(i / ((long long)N * N));
    const long long ph = ((long long)(7 * p % N) * y + (long long)(13 * p % N) * z) % N;
    double sn, cs; sincospi(2.0 * (double)ph / N, &sn, &cs);
    a[i] = make_double2(cs, sn);
}
__global__ void check_2d(const double2* a, int planes, double* maxerr)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)planes * N * N) return;
    const int z = (int)(i % N), y = (int)((i / N) % N), p = (int)(i / ((long long)N * N));
    const double ex = (y == 7 * p % N && z == 13 * p % N) ? (double)N * N : 0.0;
    const double err = fmax(fabs(a[i].x - ex), fabs(a[i].y));
    if (err > 1e-7) atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(err));
}

int main(int argc, char** argv)
{
    const int planes = argc > 1 ? atoi(argv[1]) : 512;
    const int lag_arg = argc > 2 ? atoi(argv[2]) : 0;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const long long count = (long long)planes * N * N, nlines = (long long)planes * N, ntiles = nlines / C;
    printf("device %s sms=%d  %d planes of %dx%d fp64, smem/CTA %zu B\n", prop.name, sms, planes, N, N, SMEM_BYTES);
    double2 *d_a, *d_b; double* d_err;
    CK(cudaMalloc(&d_a, count * 16)); CK(cudaMalloc(&d_b, count * 16)); CK(cudaMalloc(&d_err, 8));
    // twiddles
    std::vector<double2> lut(LUT_ENTRIES);
    for (int s = 1; s < S::NSTAGES; s++) {
        const int RAD = S::rad(s), NS = S::ns(s), off = S::lut_off(s);
        for (int m = 1; m < RAD; m++)
            for (int k = 0; k < NS; k++) {
                const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)(k * m) / (long double)(NS * RAD);
                lut[off + (m - 1) * NS + k] = make_double2((double)cosl(a), (double)sinl(a));
            }
    }
    double2* d_lut; CK(cudaMalloc(&d_lut, lut.size() * 16)); CK(cudaMemcpy(d_lut, lut.data(), lut.size() * 16, cudaMemcpyHostToDevice));
    CUtensorMap map_a = make_map(d_a, nlines), map_b = make_map(d_b, nlines);
    CK(cudaFuncSetAttribute(pass_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CK(cudaFuncSetAttribute(pass_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CK(cudaFuncSetAttribute(pass_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CK(cudaFuncSetAttribute(fused_t0_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const unsigned fb = (unsigned)((count + 255) / 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 2.0 * 16.0 * (double)count;
    double err; float ms;
    const int iters = 5;

    // Z pass
    fill_lines<<<fb, 256>>>(d_a, nlines); CK(cudaMemset(d_err, 0, 8));
    pass_tma<0><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
    CK(cudaGetLastError());
    check_lines<<<fb, 256>>>(d_b, nlines, d_err); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost));
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) pass_tma<0><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("Z pass  TMA ring : %.3f ms  %.0f GB/s  err=%.2e\n", ms, bytes / ms * 1e-6, err); fflush(stdout);

    // Y pass
    fill_cols<<<fb, 256>>>(d_a, planes); CK(cudaMemset(d_err, 0, 8));
    pass_tma<1><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
    CK(cudaGetLastError());
    check_cols<<<fb, 256>>>(d_b, planes, d_err); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost));
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) pass_tma<1><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("Y pass  TMA ring : %.3f ms  %.0f GB/s  err=%.2e\n", ms, bytes / ms * 1e-6, err); fflush(stdout);

    // X pass (needs the full cube: 512 planes)
    if (planes == N) {
        CUtensorMap map3 = make_map3(d_a, planes);
        fill_x<<<fb, 256>>>(d_a); CK(cudaMemset(d_err, 0, 8));
        pass_tma<2><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map3, map_b, d_lut, ntiles);
        CK(cudaGetLastError());
        check_x<<<fb, 256>>>(d_b, d_err); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost));
        cudaEventRecord(e0);
        for (int i = 0; i < iters; i++) pass_tma<2><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map3, map_b, d_lut, ntiles);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
        printf("X pass  TMA ring : %.3f ms  %.0f GB/s  err=%.2e\n", ms, bytes / ms * 1e-6, err); fflush(stdout);
    }

    // the LIBRARY kernels (fft_tma.cuh) in the same harness, plane-wave and random data, short and long timing loops
    {
        using LS = Sched<512, 8, 8, 8, 8>;
        CK(cudaFuncSetAttribute(fft_tma_pass_kernel<LS, double, 8, TMA_Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaGeom<LS, double, 8>::SMEM));
        CK(cudaFuncSetAttribute(fft_tma_pass_kernel<LS, double, 8, TMA_Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaGeom<LS, double, 8>::SMEM));
        CK(cudaFuncSetAttribute(fft_tma_pass_kernel<LS, double, 8, TMA_XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaGeom<LS, double, 8>::SMEM));
        const size_t lsm = TmaGeom<LS, double, 8>::SMEM;
        EncodeTiledFn fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q));
        auto enc3 = [&](void* base, int b1, int b2) {
            CUtensorMap m;
            cuuint64_t dims[3] = {(cuuint64_t)N * 2, (cuuint64_t)N, (cuuint64_t)planes};
            cuuint64_t strides[2] = {(cuuint64_t)N * 16, (cuuint64_t)N * N * 16};
            cuuint32_t box[3] = {16, (cuuint32_t)b1, (cuuint32_t)b2};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
            return m;
        };
        CUtensorMap dummy{};
        for (int rnd = 0; rnd < 2; rnd++) {
            if (rnd) { fill_random<<<fb, 256>>>(d_a, count); fill_random<<<fb, 256>>>(d_b, count); }
            else fill_lines<<<fb, 256>>>(d_a, nlines);
            CK(cudaDeviceSynchronize());
            for (int reps : {5, 40}) {
                for (int inplace = 0; inplace < 2; inplace++) {
                    double2* o = inplace ? d_a : d_b;
                    TmaArgs<double> A{}; A.lut = d_lut; A.scale = 1.0;
                    // Z
                    A.in = d_a; A.out = o; A.ntiles = ntiles; A.G = (int)ntiles; A.in_SA = A.out_SA = 0;
                    cudaEventRecord(e0);
                    for (int i = 0; i < reps; i++) fft_tma_pass_kernel<LS, double, 8, TMA_Z><<<sms, 512, lsm>>>(A, dummy, dummy);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
                    printf("LIB Z   %s data, %2d reps, %s: %.3f ms  %.0f GB/s\n", rnd ? "random" : "wave  ", reps, inplace ? "in place    " : "out of place", ms, bytes / ms * 1e-6);
                    // Y
                    CUtensorMap mi = enc3(d_a, 256, 1), mo = enc3(o, 256, 1);
                    A.G = N / 8; A.ntiles = (long long)planes * A.G;
                    cudaEventRecord(e0);
                    for (int i = 0; i < reps; i++) fft_tma_pass_kernel<LS, double, 8, TMA_Y><<<sms, 512, lsm>>>(A, mi, mo);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
                    printf("LIB Y   %s data, %2d reps, %s: %.3f ms  %.0f GB/s\n", rnd ? "random" : "wave  ", reps, inplace ? "in place    " : "out of place", ms, bytes / ms * 1e-6);
                    // X (out of place only)
                    if (!inplace && planes == N) {
                        CUtensorMap mx = enc3(d_a, 1, 256);
                        A.out = d_b; A.out_SA = (long long)N * N;
                        cudaEventRecord(e0);
                        for (int i = 0; i < reps; i++) fft_tma_pass_kernel<LS, double, 8, TMA_XF><<<sms, 512, lsm>>>(A, mx, dummy);
                        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
                        printf("LIB XF  %s data, %2d reps, out of place: %.3f ms  %.0f GB/s\n", rnd ? "random" : "wave  ", reps, ms, bytes / ms * 1e-6);
                    }
                    fflush(stdout);
                }
            }
        }
        // the prototype's own kernels on random data, long loop
        for (int reps : {5, 40}) {
            cudaEventRecord(e0);
            for (int i = 0; i < reps; i++) pass_tma<0><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            printf("PROTO Z random data, %2d reps: %.3f ms  %.0f GB/s\n", reps, ms, bytes / ms * 1e-6);
            cudaEventRecord(e0);
            for (int i = 0; i < reps; i++) pass_tma<1><<<sms, NCONS, SMEM_BYTES>>>(d_a, d_b, map_a, map_b, d_lut, ntiles);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            printf("PROTO Y random data, %2d reps: %.3f ms  %.0f GB/s\n", reps, ms, bytes / ms * 1e-6);
        }
    }
    if (getenv("TMA_T0_SKIP_FUSED")) return 0;

    // fused t0: a -> b (Z), b -> b (Y)
    unsigned long long* d_done; unsigned int* d_ticket;
    CK(cudaMalloc(&d_done, planes * 8)); CK(cudaMemset(d_done, 0, planes * 8));
    CK(cudaMalloc(&d_ticket, 8)); CK(cudaMemset(d_ticket, 0, 8));
    FusedT0 F{d_a, d_b, d_done, d_ticket, 0, planes, 0};
    const int G = N / C;
    int lags[4] = {lag_arg > 0 ? lag_arg : (sms * NSLOT + 2 * G - 1) / (2 * G) + 1, 3, 6, 10};
    unsigned long long epoch = 0;
    for (int li = 0; li < (lag_arg > 0 ? 1 : 4); li++) {
        F.lag = lags[li];
        fill_2d<<<fb, 256>>>(d_a, planes); CK(cudaMemset(d_err, 0, 8));
        F.target = ++epoch * (unsigned long long)G;
        fused_t0_tma<<<sms, NCONS + 32, SMEM_BYTES>>>(F, map_b, d_lut);
        CK(cudaGetLastError());
        check_2d<<<fb, 256>>>(d_b, planes, d_err); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost));
        cudaEventRecord(e0);
        for (int i = 0; i < iters; i++) {
            F.target = ++epoch * (unsigned long long)G;
            fused_t0_tma<<<sms, NCONS + 32, SMEM_BYTES>>>(F, map_b, d_lut);
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
        printf("fused t0 TMA ring (lag %2d): %.3f ms  (2-sweep equivalent %.0f GB/s, compulsory-traffic %.0f GB/s)  err=%.2e\n", F.lag, ms,
               2 * bytes / ms * 1e-6, bytes / ms * 1e-6, err);
        fflush(stdout);
    }
    return 0;
}
