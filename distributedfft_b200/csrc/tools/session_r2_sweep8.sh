#!/bin/bash
# round-2 8-GPU tuning sweep (developer tool): one torchrun, every variant in-process (see sweep.py)
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 distributedfft_b200/csrc/tools/sweep.py \
  "512:double:0:nopipe" "512:double:5:nopipe" \
  "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1;DFFT_SIGNAL_KERNELS=1" \
  "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=0" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=2" \
  "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=2;DFFT_PIPE_CAP=1" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=8;DFFT_PIPE_CAP=1" \
  "512:double:5;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" "512:double:0:nofuse;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" \
  "512:double:0:nccl+nopipe" "512:double:0:nccl;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" "512:double:0:nccl;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=0" \
  "1024:double:0:nopipe" "1024:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" "1024:double:0;DFFT_PIPELINE=1;DFFT_PARTS=8;DFFT_PIPE_CAP=1" \
  "768:float:0:nopipe" "768:float:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" \
  "512:double:0:nopipe" "512:double:0;DFFT_PIPELINE=1;DFFT_PARTS=4;DFFT_PIPE_CAP=1" \
  2>&1 | grep -v "^W\|Warn\|warn\|^\*" | tee gpurun_out/r2_sweep_n$N.log | tail -24
