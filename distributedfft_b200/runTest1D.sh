#!/bin/bash
# Batched 1-D sweep with the surface of the reference's templateFFT/batchTest/runTest1D_opt.sh (same lengths, same CSV
# header, 2^26 points per run): geometric ladders of 2 (from 256), 3, 5 and 7 up to the reference's limits.
# Lengths beyond one shared-memory line (6400 points in double) run on the two-pass four-step plan;
# what is still unsupported is reported and skipped.
here="$(cd "$(dirname "$0")" && pwd)"
iters=${NUM_ITER:-1000}
csv=${CSV:-batch_result1D.csv}
echo 'X,Y,Z,Buffer,hip_time,GFlops,num_iter,bandwidth,max error' > "$csv"

ladder() {   # ladder <base> <first> <last>
  local x=$2
  while [ "$x" -le "$3" ]; do
    "$here/batchFFT" 1d "$x" 1 1 "$iters" 0 "$csv" || echo "length $x: skipped (unsupported)"
    x=$((x * $1))
  done
}
ladder 2 256 131072
ladder 3 3 14348907
ladder 5 5 48828125
ladder 7 7 40353607
