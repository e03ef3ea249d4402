#!/bin/bash
# One GPU call that refreshes the evidence under profiles/: other shapes, the launch list of the bench command and one
# `ncu --set full` capture per hot kernel (each after the same command has exited 0 without ncu).
set -x
cd "$(dirname "$0")/.."
python scripts/dev_c4_check.py > gpurun_out/c4_check.log 2>&1
python scripts/dev_c5_check.py 100000 1024 > gpurun_out/c5_check.log 2>&1
python scripts/dev_c5_check.py 12800 128 > gpurun_out/c5_check_128.log 2>&1
python scripts/dev_project_check.py 4096 > gpurun_out/project_check.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --device-leg-only > gpurun_out/bench_device_leg.json 2> gpurun_out/bench_device_leg.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v4.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --device-leg-only > gpurun_out/ncu_launch_v4.log 2>&1
python scripts/prof_k2.py 128 2 > gpurun_out/prof_plain_v4.log 2>&1 || exit 1
for k in decode_kernel cdc_kernel cov_syrk_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_v4 \
      python scripts/prof_k2.py 128 2 > gpurun_out/ncu_${k}_v4.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
