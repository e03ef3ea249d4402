#!/bin/bash
# round 2, GPU call 1: correctness of the new engine + first numbers
mkdir -p gpurun_out/r2
{ nproc; free -g | head -2; df -h /dev/shm | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nvidia-smi topo -m; } > gpurun_out/r2/box_facts.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log
tail -5 gpurun_out/r2/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput.log 2>&1; echo "cli rc=$?"
tail -30 gpurun_out/r2/cli_throughput.log
