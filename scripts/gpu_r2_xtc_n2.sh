#!/bin/bash
# round 2: the bench under torchrun with 2 ranks after the e2e_xtc leg was added (its agreement + all-reduce path)
O=gpurun_out/r2/xtc_n2; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-cli > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -4 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2/xtc_n2/bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "trr", d["e2e_trr"]["value"], "allreduce", d["allreduce_check"]["ok"] if d.get("allreduce_check") else None)
print("xtc", json.dumps(d.get("e2e_xtc"))[:900])
PY
