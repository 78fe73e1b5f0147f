#!/bin/bash
# round-2 final single-GPU validation + measurement session (run under gpurun from the repo root): the driver's own checks
# (pytest -m gpu, smoke, both bench arms) at HEAD, then the ncu evidence for profiles/.
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > $O/r2_final_pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_final_pytest_1gpu.log
timeout 200 python __graft_entry__.py smoke > $O/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2_final_smoke.log
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_final_bench_ref.json 2> $O/r2_final_bench_ref.err; echo "ref rc=$?"; cut -c1-400 $O/r2_final_bench_ref.json
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; echo "bench rc=$?"; cut -c1-2200 $O/r2_final_bench_n1.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/r2_final_clocks_idle.csv
# launch list of the bench command (shares, not absolutes: ncu serialises and runs cold)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 60 --csv --log-file $O/r2_final_launches_bench_n1.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $O/r2_final_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# the three pass kernels of one timed step, full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fft_tma_pass_kernel|fft_tile_kernel' -s 9 -c 3 -o $O/r2_final_prof_512 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r2_final_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/r2_final_prof_512.ncu-rep
# the reference driver's surface and the batched 1-D / 2-D benchmark surface (Test_1D / Test_2D CSV columns)
timeout 120 distributedfft_b200/distFFT 512 512 512 1 > $O/r2_final_driver_512_1gpu.log 2>&1; tail -12 $O/r2_final_driver_512_1gpu.log
NUM_ITER=20 CSV=$O/r2_final_batch1d.csv timeout 400 bash distributedfft_b200/runTest1D.sh 2>&1 | grep -E "^FFT:|skipped" | tail -60 > $O/r2_final_batch1d.log; tail -3 $O/r2_final_batch1d.log
NUM_ITER=20 CSV=$O/r2_final_batch2d.csv timeout 400 bash distributedfft_b200/runTest2D.sh 2>&1 | grep -E "^FFT:|skipped" | tail -40 > $O/r2_final_batch2d.log; tail -3 $O/r2_final_batch2d.log
