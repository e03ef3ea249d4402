#!/bin/bash
# round 2, GPU call 4: K3 v4 (two wide TMA boxes per stage), K3 buffers sized up front
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q -k "covariance or c5 or engine or pypca or two_rank or zero_frames or writers or cli" > gpurun_out/r2/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu5.log
tail -6 gpurun_out/r2/pytest_gpu5.log
timeout 300 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench5_C5.json 2> gpurun_out/r2/bench5_C5.err; echo "bench C5 rc=$?"
timeout 300 python bench.py --workload C5 --frames-per-step 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench5_C5_f256.json 2> gpurun_out/r2/bench5_C5_f256.err; echo "bench C5 f256 rc=$?"
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput5.log 2>&1; echo "cli rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2/c5_launches5.csv python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_c5_list5.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cov_syrk -s 2 -c 1 -f -o gpurun_out/r2/k3_v5 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_k3_5.log 2>&1; echo "ncu k3 rc=$?"
