#!/bin/bash
# round 2, GPU call 3: K2 grid-rounded atomics, K3 v3 (128x128 tiles), cached launch queries
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu3.log
tail -6 gpurun_out/r2/pytest_gpu3.log
timeout 300 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench3_C5.json 2> gpurun_out/r2/bench3_C5.err; echo "bench C5 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench3_c2.json 2> gpurun_out/r2/bench3_c2.err; echo "bench C2 rc=$?"
tail -3 gpurun_out/r2/bench3_c2.err
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline --no-cli > gpurun_out/r2/bench3_C4.json 2> gpurun_out/r2/bench3_C4.err; echo "bench C4 rc=$?"
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput3.log 2>&1; echo "cli rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2/c5_launches.csv python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_c5_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cov_syrk -s 2 -c 1 -f -o gpurun_out/r2/k3_v3 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
