#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu8.log
tail -5 gpurun_out/r2/pytest_gpu8.log
timeout 300 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench8_C5.json 2> gpurun_out/r2/bench8_C5.err; echo "bench C5 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench8_c2.json 2> gpurun_out/r2/bench8_c2.err; echo "bench C2 rc=$?"
tail -3 gpurun_out/r2/bench8_c2.err
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput8.log 2>&1; echo "cli rc=$?"
grep "cgc trace" gpurun_out/r2/cli_throughput8.log | cut -c1-460
