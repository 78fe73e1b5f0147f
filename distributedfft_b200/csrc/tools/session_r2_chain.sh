#!/bin/bash
# round-2: the kernel-chain pipeline (fft_fused_yx_kernel, role-pinned) against the plain path
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 200 -k "vs_oracle and (128-128-128 or 1024-1024) and (pipe or p2p)" 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 distributedfft_b200/csrc/tools/sweep.py \
  "512:double:0:nopipe" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4" \
  "1024:double:0:nopipe" "1024:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2" "1024:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4" \
  "768:float:0:nopipe" "768:float:0;DFFT_PIPELINE=1;DFFT_PARTS=2" "256:double:0:nopipe" "256:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2" \
  "512:double:0:nopipe" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4" \
  2>&1 | grep -v "^W\|Warn\|warn\|^\*\|OMP_NUM" | tee gpurun_out/r2_chain_pinned_n$N.log | tail -16
