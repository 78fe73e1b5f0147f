#!/bin/bash
# Batched 2-D sweep with the surface of the reference's templateFFT/batchTest/runTest2D_opt.sh: every (X, Y) pair of
# {2048, 1024, 512, 256, 128}, 2^26 points per run, same CSV header.
here="$(cd "$(dirname "$0")" && pwd)"
iters=${NUM_ITER:-1000}
csv=${CSV:-batch_result2D.csv}
echo 'X,Y,Z,Buffer,hip_time,GFlops,num_iter,bandwidth,max error' > "$csv"
sizes="2048 1024 512 256 128"
for x in $sizes; do
  for y in $sizes; do
    "$here/batchFFT" 2d "$x" "$y" 1 "$iters" 0 "$csv" || echo "${x}x${y}: skipped"
  done
done
