#!/bin/bash
# round 2, second 8-GPU call: bench at N = 8 and 4 with device placement and link-rate-weighted e2e sharding
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out/r2
for n in 8 4; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2/bench_multi3_n$n.json 2> gpurun_out/r2/bench_multi3_n$n.err; echo "bench n=$n rc=$?"
    tail -2 gpurun_out/r2/bench_multi3_n$n.err
  fi
done
timeout 300 python -m pytest tests/test_engine.py -m gpu -x -q -k "nccl or two_gpus" > gpurun_out/r2/pytest_multi3.log 2>&1; tail -2 gpurun_out/r2/pytest_multi2.log
