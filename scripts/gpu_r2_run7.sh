#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu7.log
tail -5 gpurun_out/r2/pytest_gpu7.log
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput7.log 2>&1; echo "cli rc=$?"
grep "cgc trace" gpurun_out/r2/cli_throughput7.log | cut -c1-560
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench7_c2.json 2> gpurun_out/r2/bench7_c2.err; echo "bench C2 rc=$?"
bash scripts/gpu_sanitizer.sh
