#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): the multi-GPU tests, the H2D matrix, bench at every N up to the box's GPU count
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out/r2
{ nproc; free -g | head -2; nvidia-smi topo -m; ls /sys/devices/system/node/ 2>/dev/null; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)"; } > gpurun_out/r2/box_facts_n$N.txt 2>&1
timeout 300 scripts/h2d_matrix 512 4 > gpurun_out/r2/h2d_matrix_n$N.txt 2>&1; echo "h2d matrix rc=$?"
tail -12 gpurun_out/r2/h2d_matrix_n$N.txt
timeout 600 python -m pytest tests/test_engine.py tests/test_dist.py -m gpu -x -q > gpurun_out/r2/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2/pytest_multi_n$N.log
if [ $N -ge 8 ]; then
  timeout 600 python scripts/dev_cli_throughput.py 8192 /dev/shm 8 > gpurun_out/r2/cli_throughput_n8.log 2>&1; echo "cli n8 rc=$?"
  grep -E "cgc trace|^gpus" gpurun_out/r2/cli_throughput_n8.log | cut -c1-300 | tail -30
fi
for n in 8 4; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2/bench_multi_n$n.json 2> gpurun_out/r2/bench_multi_n$n.err; echo "bench n=$n rc=$?"
    tail -2 gpurun_out/r2/bench_multi_n$n.err
  fi
done
