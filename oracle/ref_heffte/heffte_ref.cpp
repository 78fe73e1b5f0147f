// heffte_ref.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).  C entry points around the UNMODIFIED heFFTe 2.1.0 vendored
// in the reference tree (heffte/heffteBenchmark), built with its dependency-free `stock` CPU backend: the library
// the reference benchmarks itself against (heffte/heffteBenchmark/benchmarks/speed3d.h, heffteSpeed.sh:
// `speed3d_c2c <backend> double X Y Z -slabs -p2p_pl`) and the only CPU FFT in the tree.  It is used
//   * to pin the repo's restated oracle: the forward / backward spectrum of the same world array computed by the
//     reference's own code over P slab "ranks" (threads behind the mpi.h stand-in), and
//   * as the `--impl reference` CPU arm of bench.py (kind "reference").
// World arrays are natural order A[x][y][z], z fastest (3dmpifft_opt/fftSpeed3d_c2c.cpp:56-62); heFFTe's index 0 is
// the fastest one, so its box is (z, y, x).  Input boxes are x-slabs, output boxes y-slabs, like the reference's
// slab decomposition (3dmpifft_opt/include/fft_mpi_3d_api.cpp:56-66, 535-536) and speed3d's `-slabs`.
#include <algorithm>
#include <complex>
#include <cstring>
#include <vector>

#include "heffte.h"

namespace {

struct Job {
    int n0, n1, n2, P, direction, algorithm, reps, warmup, scale_full, pair_reps;
    const void* in;
    void* out;
    double* times;      // [reps] forward seconds (max over ranks), rank 0 writes
    double* pair_time;  // speed3d protocol: mean of (forward + backward) / 2
    int rc;
};

heffte::plan_options make_options(int algorithm)
{
    heffte::plan_options o = heffte::default_options<heffte::backend::stock>();
    o.use_pencils = false;   // -slabs
    switch (algorithm) {
        case 1: o.algorithm = heffte::reshape_algorithm::alltoall; break;
        case 2: o.algorithm = heffte::reshape_algorithm::p2p_plined; break;   // -p2p_pl (heffteSpeed.sh)
        case 3: o.algorithm = heffte::reshape_algorithm::p2p; break;
        default: o.algorithm = heffte::reshape_algorithm::alltoallv; break;
    }
    return o;
}

template <typename T> void transform_rank(void* vjob)
{
    Job& j = *static_cast<Job*>(vjob);
    using cplx = std::complex<T>;
    MPI_Comm comm = MPI_COMM_WORLD;
    const int me = heffte::mpi::comm_rank(comm);
    heffte::box3d<> const world = {{0, 0, 0}, {j.n2 - 1, j.n1 - 1, j.n0 - 1}};
    std::vector<heffte::box3d<>> inboxes = heffte::split_world(world, {1, 1, j.P});    // x-slabs
    std::vector<heffte::box3d<>> outboxes = heffte::split_world(world, {1, j.P, 1});   // y-slabs
    // the backward transform goes from the y-slab boxes back to the x-slab boxes
    heffte::fft3d<heffte::backend::stock> fft(inboxes[me], outboxes[me], comm, make_options(j.algorithm));
    const heffte::box3d<>& ib = j.direction > 0 ? inboxes[me] : outboxes[me];
    const heffte::box3d<>& ob = j.direction > 0 ? outboxes[me] : inboxes[me];
    std::vector<cplx> a((size_t)std::max(fft.size_inbox(), fft.size_outbox())), b(a.size());
    const cplx* win = static_cast<const cplx*>(j.in);
    cplx* wout = static_cast<cplx*>(j.out);
    auto widx = [&](long long x, long long y, long long z) { return ((size_t)x * j.n1 + y) * j.n2 + z; };
    size_t k = 0;
    for (int x = ib.low[2]; x <= ib.high[2]; x++)
        for (int y = ib.low[1]; y <= ib.high[1]; y++)
            for (int z = ib.low[0]; z <= ib.high[0]; z++) a[k++] = win[widx(x, y, z)];
    if (j.direction > 0) fft.forward(a.data(), b.data());
    else fft.backward(a.data(), b.data(), j.scale_full ? heffte::scale::full : heffte::scale::none);
    k = 0;
    for (int x = ob.low[2]; x <= ob.high[2]; x++)
        for (int y = ob.low[1]; y <= ob.high[1]; y++)
            for (int z = ob.low[0]; z <= ob.high[0]; z++) wout[widx(x, y, z)] = b[k++];
}

// speed3d's timing loop (benchmarks/speed3d.h:96-118) plus a forward-only loop (the metric bench.py quotes)
template <typename T> void time_rank(void* vjob)
{
    Job& j = *static_cast<Job*>(vjob);
    using cplx = std::complex<T>;
    MPI_Comm comm = MPI_COMM_WORLD;
    const int me = heffte::mpi::comm_rank(comm);
    heffte::box3d<> const world = {{0, 0, 0}, {j.n2 - 1, j.n1 - 1, j.n0 - 1}};
    std::vector<heffte::box3d<>> inboxes = heffte::split_world(world, {1, 1, j.P});
    std::vector<heffte::box3d<>> outboxes = heffte::split_world(world, {1, j.P, 1});
    heffte::fft3d<heffte::backend::stock> fft(inboxes[me], outboxes[me], comm, make_options(j.algorithm));
    std::vector<cplx> data((size_t)std::max(fft.size_inbox(), fft.size_outbox()));
    unsigned long long s = 4242ull + 977ull * me;   // values do not affect FFT time; U(0,1) like test_fft3d.h:20-27
    for (auto& v : data) { s = s * 48271ull % 2147483647ull; v = cplx((T)((double)s / 2147483647.0), 0); }
    std::vector<cplx> work(fft.size_workspace());
    for (int w = 0; w < j.warmup; w++) {
        fft.forward(data.data(), data.data(), work.data(), heffte::scale::full);
        fft.backward(data.data(), data.data(), work.data());
    }
    for (int r = 0; r < j.reps; r++) {   // forward only
        MPI_Barrier(comm);
        double t = -MPI_Wtime();
        fft.forward(data.data(), data.data(), work.data());
        MPI_Barrier(comm);
        t += MPI_Wtime();
        if (me == 0 && j.times) j.times[r] = t;
        fft.backward(data.data(), data.data(), work.data(), heffte::scale::full);   // keeps the values bounded
    }
    if (j.pair_reps < 1) return;
    MPI_Barrier(comm);
    double t = -MPI_Wtime();
    for (int r = 0; r < j.pair_reps; r++) {
        fft.forward(data.data(), data.data(), work.data(), heffte::scale::full);
        fft.backward(data.data(), data.data(), work.data());
    }
    MPI_Barrier(comm);
    t += MPI_Wtime();
    if (me == 0 && j.pair_time) *j.pair_time = t / (2.0 * j.pair_reps);
}

}  // namespace

extern "C" {

int heffte_ref_version(void) { return Heffte_VERSION_MAJOR * 100 + Heffte_VERSION_MINOR * 10 + Heffte_VERSION_PATCH; }

/* world-array transform by the reference library over P slab ranks.  precision 0 = double, 1 = float (interleaved complex);
 * direction +1 forward (e^{-i...}, unnormalised), -1 backward (unnormalised unless scale_full); algorithm 0 alltoallv,
 * 1 alltoall, 2 p2p_plined, 3 p2p.  Returns 0, or -1 when P does not fit the box. */
int heffte_ref_fft3d_c2c(int n0, int n1, int n2, int P, const void* in, void* out, int direction, int algorithm, int precision, int scale_full)
{
    if (P < 1 || P > n0 || P > n1 || n0 < 1 || n1 < 1 || n2 < 1) return -1;
    Job j{n0, n1, n2, P, direction, algorithm, 0, 0, scale_full, 0, in, out, nullptr, nullptr, 0};
    try {
        if (precision == 0) tmpi_run(P, transform_rank<double>, &j);
        else tmpi_run(P, transform_rank<float>, &j);
    } catch (...) { return -2; }
    return 0;
}

/* times[reps] = forward seconds per repetition; *pair_time = speed3d's figure, mean of (forward + backward) / 2 over pair_reps (0: skip) */
int heffte_ref_time(int n0, int n1, int n2, int P, int reps, int warmup, int pair_reps, int algorithm, int precision, double* times, double* pair_time)
{
    if (P < 1 || P > n0 || P > n1 || reps < 1) return -1;
    Job j{n0, n1, n2, P, 1, algorithm, reps, warmup, 0, pair_reps, nullptr, nullptr, times, pair_time, 0};
    try {
        if (precision == 0) tmpi_run(P, time_rank<double>, &j);
        else tmpi_run(P, time_rank<float>, &j);
    } catch (...) { return -2; }
    return 0;
}
}
