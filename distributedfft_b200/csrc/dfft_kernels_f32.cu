// fp32 instantiations.  Per-length choices come from kbench sweeps on B200 (profiles/kbench*.log):
// small CTAs (finer barriers) win for the contiguous pass, 128-byte row segments (C = 8 complex
// doubles) for the strided ones.
#include "dfft_kernels_inst.cuh"
namespace dfft {
void register_f32(std::vector<SizeEntry>& v)
{
    using T = float;
    v.push_back(make_entry<T, Sched<4, 2, 2, 2>, 64, false, 4, true, Sched<4, 2, 2, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<8, 4, 4, 2>, 64, false, 4, true, Sched<8, 4, 4, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<16, 4, 4, 4>, 32, false, 4, true, Sched<16, 4, 4, 4>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<32, 8, 8, 4>, 32, false, 4, true, Sched<32, 8, 8, 4>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<64, 8, 8, 8>, 16, true, 4, true, Sched<64, 8, 8, 8>, 16, true, 4, true>());
    v.push_back(make_entry<T, Sched<128, 16, 16, 8>, 16, false, 4, true, Sched<128, 16, 16, 8>, 16, false, 4, true>());
    v.push_back(make_entry<T, Sched<256, 16, 16, 16>, 8, false, 4, true, Sched<256, 16, 16, 16>, 8, false, 4, true>());
    v.push_back(make_entry<T, Sched<512, 8, 8, 8, 8>, 2, true, 4, true, Sched<512, 16, 8, 8, 8>, 8, false, 2, false>());
    v.push_back(make_entry<T, Sched<1024, 16, 16, 8, 8>, 2, false, 4, true, Sched<1024, 16, 16, 8, 8>, 4, false, 2, false>());
    v.push_back(make_entry<T, Sched<2048, 16, 16, 16, 8>, 1, false, 4, false, Sched<2048, 16, 16, 16, 8>, 4, false, 1, false>());
    v.push_back(make_entry<T, Sched<4096, 16, 16, 16, 16>, 1, false, 2, false, Sched<4096, 16, 16, 16, 16>, 2, false, 1, false>());
    // mixed radix
    v.push_back(make_entry<T, Sched<768, 12, 4, 4, 4, 4, 3>, 2, false, 4, true, Sched<768, 12, 4, 4, 4, 4, 3>, 4, false, 2, false>());
    v.push_back(make_entry<T, Sched<384, 12, 4, 4, 4, 3, 2>, 4, false, 4, true, Sched<384, 12, 4, 4, 4, 3, 2>, 8, false, 2, false>());
    v.push_back(make_entry<T, Sched<192, 12, 4, 4, 4, 3>, 8, false, 4, true, Sched<192, 12, 4, 4, 4, 3>, 8, false, 4, true>());
    v.push_back(make_entry<T, Sched<96, 12, 4, 4, 3, 2>, 16, false, 4, true, Sched<96, 12, 4, 4, 3, 2>, 16, false, 4, true>());
    v.push_back(make_entry<T, Sched<48, 12, 4, 4, 3>, 32, false, 4, true, Sched<48, 12, 4, 4, 3>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<24, 6, 3, 2, 2, 2>, 32, false, 4, true, Sched<24, 6, 3, 2, 2, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<12, 6, 3, 2, 2>, 64, false, 4, true, Sched<12, 6, 3, 2, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<6, 6, 3, 2>, 64, false, 4, true, Sched<6, 6, 3, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<9, 3, 3, 3>, 32, false, 4, true, Sched<9, 3, 3, 3>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<10, 10, 5, 2>, 64, false, 4, true, Sched<10, 10, 5, 2>, 32, false, 4, true>());
    v.push_back(make_entry<T, Sched<100, 10, 5, 5, 2, 2>, 16, false, 4, true, Sched<100, 10, 5, 5, 2, 2>, 16, false, 4, true>());
    v.push_back(make_entry<T, Sched<125, 5, 5, 5, 5>, 8, false, 4, true, Sched<125, 5, 5, 5, 5>, 8, false, 4, true>());
    v.push_back(make_entry<T, Sched<49, 7, 7, 7>, 16, false, 4, true, Sched<49, 7, 7, 7>, 16, false, 4, true>());
    v.push_back(make_entry<T, Sched<243, 9, 3, 3, 3, 3, 3>, 8, false, 4, true, Sched<243, 9, 3, 3, 3, 3, 3>, 8, false, 4, true>());
    v.push_back(make_entry<T, Sched<1000, 10, 5, 5, 5, 2, 2, 2>, 2, false, 4, true, Sched<1000, 10, 5, 5, 5, 2, 2, 2>, 4, false, 2, false>());
}
}  // namespace dfft
