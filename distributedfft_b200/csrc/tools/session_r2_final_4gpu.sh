#!/bin/bash
# round-2 final 4-GPU check: BASELINE config C3 at full size (plain and kernel-chain schedules) and the bench lines at N = 4
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q --timeout 300 -k "C3 or C3-chain" > $O/r2_final_pytest_4gpu_configs.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_final_pytest_4gpu_configs.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 2> $O/r2_final_bench_n4.err | grep "^{" > $O/r2_final_bench_n4.json; echo "bench rc=$?"; cut -c1-1500 $O/r2_final_bench_n4.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 --size 1024 --no-e2e 2>/dev/null | grep "^{" > $O/r2_final_bench_n4_1024.json; cut -c1-900 $O/r2_final_bench_n4_1024.json
