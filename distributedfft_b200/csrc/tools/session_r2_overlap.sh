#!/bin/bash
# round-2: the single-kernel overlapped forward (fft_fused3_kernel, reworked signalling) against the default path
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
N=${1:-2}
DFFT_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 200 -k "overlapped" 2>&1 | tail -15
DFFT_DEBUG_TIMELINE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 distributedfft_b200/csrc/tools/sweep.py \
  "512:double:0:nopipe" "512:double:0:overlap" "512:double:0:overlap;DFFT_PARTS=2" "512:double:0:overlap;DFFT_PARTS=8" \
  "512:double:0:overlap;DFFT_PARTS=1" "512:double:0:overlap;DFFT_LAG=3" "512:double:0:overlap;DFFT_LAG=8" \
  "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4" \
  "1024:double:0:nopipe" "1024:double:0:overlap" "768:float:0:nopipe" "768:float:0:overlap" \
  "512:double:0:nopipe" "512:double:0:overlap" \
  2>&1 | grep -v "^W\|Warn\|warn\|^\*\|OMP_NUM" | tee gpurun_out/r2_overlap_n$N.log | tail -20
