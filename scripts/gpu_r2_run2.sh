#!/bin/bash
# round 2, GPU call 2: new K3 + full-size parity + bench (all workloads) + CLI
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu2.log
tail -15 gpurun_out/r2/pytest_gpu2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench2_c2.json 2> gpurun_out/r2/bench2_c2.err; echo "bench C2 rc=$?"
tail -3 gpurun_out/r2/bench2_c2.err
for w in C5 C1 C4; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-cli > gpurun_out/r2/bench2_$w.json 2> gpurun_out/r2/bench2_$w.err; echo "bench $w rc=$?"
  tail -2 gpurun_out/r2/bench2_$w.err
done
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput2.log 2>&1; echo "cli rc=$?"
