#!/bin/bash
# speedTest.sh <gpus> <X> <Y> <Z>   -- same surface as 3dmpifft_opt/speedTest.sh:1-9
# (the reference starts <ranks> MPI ranks x 1 GPU; here one process drives <gpus> GPUs over NVLink)
DIR="$(cd "$(dirname "$0")" && pwd)"
"$DIR/distFFT" "$2" "$3" "$4" "$1"
