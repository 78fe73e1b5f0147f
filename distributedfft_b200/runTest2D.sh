#!/bin/bash
# Sweep of batched 2-D transforms with the surface of templateFFT/batchTest/runTest2D_opt.sh:1-12
# (X, Y from 2048 down to 128; 2^26 points per run, CSV batch_result2D.csv).
DIR="$(cd "$(dirname "$0")" && pwd)"
num_iter=${NUM_ITER:-1000}
printResult=0
csv=${CSV:-batch_result2D.csv}
echo 'X,Y,Z,Buffer,hip_time,GFlops,num_iter,bandwidth,max error' > "$csv"
for ((X=2048; X>=128; X=X/2)); do
  for ((Y=2048; Y>=128; Y=Y/2)); do
    "$DIR/batchFFT" 2d $X $Y 1 "$num_iter" "$printResult" "$csv" || echo "X=$X Y=$Y: skipped"
  done
done
