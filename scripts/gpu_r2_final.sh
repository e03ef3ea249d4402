#!/bin/bash
# round 2, final one-GPU evidence run: tests, the bench lines of every workload, the reference arm, ncu captures, CLI at 10k frames
O=gpurun_out/r2/final; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python -m pytest tests/test_full_size.py -m gpu -q -s -k "full_frame" > $O/full_size_parity.log 2>&1; grep -E "max scaled|passed|failed" $O/full_size_parity.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref arm rc=$?"
timeout 600 python bench.py --workload C1 --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "C1 rc=$?"
timeout 600 python bench.py --workload C1 --impl reference --steps 20 --warmup 3 > $O/bench_c1_reference_arm.json 2> $O/bench_c1_reference_arm.err; echo "C1 ref rc=$?"
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; echo "C4 rc=$?"
timeout 600 python bench.py --workload C5 --steps 20 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "C5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --device-leg-only > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cdc_kernel -s 4 -c 1 -f -o $O/k2_cdc python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-leg-only > $O/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $O/k1_decode python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-leg-only > $O/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cov_syrk -s 2 -c 1 -f -o $O/k3_syrk python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cov_centre -s 2 -c 1 -f -o $O/k3_centre python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_k3c.log 2>&1; echo "ncu k3c rc=$?"
timeout 600 python scripts/dev_cli_throughput.py 10000 /dev/shm 1 > $O/cli_throughput_10k.log 2>&1; echo "cli 10k rc=$?"
grep -E "^gpus 1" $O/cli_throughput_10k.log | cut -c1-420
