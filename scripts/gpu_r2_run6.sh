#!/bin/bash
mkdir -p gpurun_out/r2
scripts/pinned_alloc_bench 128 > gpurun_out/r2/pinned_alloc.txt 2>&1; cat gpurun_out/r2/pinned_alloc.txt
timeout 400 python scripts/dev_cli_throughput.py 2048 /dev/shm 1 > gpurun_out/r2/cli_throughput6.log 2>&1; echo "cli rc=$?"
grep "cgc trace" gpurun_out/r2/cli_throughput6.log | cut -c1-560
