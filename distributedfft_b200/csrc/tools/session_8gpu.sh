#!/bin/bash
# Developer tool: one compact 8-GPU measurement session (gpurun --gpus 8; charged 8x, keep it short).
mkdir -p gpurun_out
T=distributedfft_b200/csrc/tools
B="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29600"
DFFT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "overlapped or (8 and vs_oracle) or spectrum" 2>&1 | tail -8 > gpurun_out/s8_pytest.log; tail -4 gpurun_out/s8_pytest.log
for n in 8 4 2; do timeout 200 $B --nproc-per-node $n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu 2>/dev/null | tee gpurun_out/s8_n$n.json | python $T/brief.py "512 n=$n"; done
timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --overlap 2>/dev/null | tee gpurun_out/s8_n8_overlap.json | python $T/brief.py "512 n=8 overlap"
for k in 2 8; do DFFT_PARTS=$k timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --overlap 2>/dev/null | python $T/brief.py "512 n=8 overlap parts=$k"; done
DFFT_VARIANT=5 timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python $T/brief.py "512 n=8 peer rows 256B"
timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --size 1024 --no-e2e --no-cpu 2>/dev/null | tee gpurun_out/s8_1024.json | python $T/brief.py "1024 n=8"
timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --size 1024 --no-e2e --no-cpu --overlap 2>/dev/null | python $T/brief.py "1024 n=8 overlap"
timeout 200 $B --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --size 768 --precision float --no-cpu 2>/dev/null | tee gpurun_out/s8_768f.json | python $T/brief.py "768f n=8"
