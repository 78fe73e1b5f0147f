#!/bin/bash
# Developer tool: one compact single-GPU measurement session (run under gpurun from the repo root).
# Output goes to gpurun_out/.  Covers: the GPU test suite incl. the gated experimental paths, the queued fused-t0
# experiments (DESIGN.md "Queued experiments"), and the headline bench line.
mkdir -p gpurun_out
T=distributedfft_b200/csrc/tools
DFFT_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/s1_pytest.log; tail -6 gpurun_out/s1_pytest.log
{
  timeout 120 python $T/sweep.py 512:double:0 512:double:0:fuse
  for v in 1 2 3 4 6 7 8; do timeout 60 python $T/sweep.py 512:double:$v:fuse; done
  for lag in 2 4; do DFFT_LAG=$lag timeout 60 python $T/sweep.py 512:double:3:fuse | sed "s/^/lag=$lag /"; done
} 2>&1 | tee gpurun_out/s1_fused_variants.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
python $T/brief.py bench < gpurun_out/s1_bench.json
DFFT_EXPERIMENTAL_LONG=1 NUM_ITER=50 CSV=gpurun_out/s1_batch1d.csv timeout 300 bash distributedfft_b200/runTest1D.sh 2>&1 | grep -E "^FFT:|skipped" | tail -40 > gpurun_out/s1_batch1d.log
