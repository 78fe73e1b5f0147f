#!/bin/bash
# Sweep of batched 1-D transforms with the surface of templateFFT/batchTest/runTest1D_opt.sh:1-21
# (powers of 2 from 256, powers of 3, 5, 7; 2^26 points per run, CSV batch_result1D.csv).  Lengths beyond one
# shared-memory line (6400 points in double) are reported as unsupported and skipped: the reference handles
# them with multi-upload passes (templateFFT.cpp:4007-4106), this library does not yet.
DIR="$(cd "$(dirname "$0")" && pwd)"
num_iter=${NUM_ITER:-1000}
printResult=0
csv=${CSV:-batch_result1D.csv}
echo 'X,Y,Z,Buffer,hip_time,GFlops,num_iter,bandwidth,max error' > "$csv"
run() { "$DIR/batchFFT" 1d "$1" 1 1 "$num_iter" "$printResult" "$csv" || echo "X=$1: skipped (unsupported length)"; }
for ((X=256; X<=131072; X=X*2)); do run $X; done
for ((X=3; X<=14348907; X=X*3)); do run $X; done
for ((X=5; X<=48828125; X=X*5)); do run $X; done
for ((X=7; X<=40353607; X=X*7)); do run $X; done
