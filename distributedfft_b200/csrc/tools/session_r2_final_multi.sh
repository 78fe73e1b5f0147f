#!/bin/bash
# round-2 multi-GPU validation + measurement session: session_r2_final_multi.sh <gpus>  (run under gpurun --gpus N)
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
N=${1:-8}
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_00_smoke.py tests/test_gpu_baseline_configs.py tests/test_gpu_multi.py -m gpu -x -q --timeout 400 > $O/r2_final_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_final_pytest_${N}gpu.log
for n in $N $((N/2)); do
  [ $n -ge 2 ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 50 --warmup 5 2> $O/r2_final_bench_n$n.err | grep "^{" > $O/r2_final_bench_n$n.json; echo "bench n=$n rc=$?"; cut -c1-1500 $O/r2_final_bench_n$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 --size 1024 --no-e2e 2>/dev/null | grep "^{" > $O/r2_final_bench_n${N}_1024.json; cut -c1-900 $O/r2_final_bench_n${N}_1024.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --size 768 --precision float --no-e2e 2>/dev/null | grep "^{" > $O/r2_final_bench_n${N}_768f.json; cut -c1-900 $O/r2_final_bench_n${N}_768f.json
timeout 300 distributedfft_b200/distFFT 512 512 512 $N > $O/r2_final_driver_512_${N}gpu.log 2>&1; tail -14 $O/r2_final_driver_512_${N}gpu.log
