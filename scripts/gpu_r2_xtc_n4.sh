#!/bin/bash
# round 2: the bench under torchrun with 4 ranks at the final commit (e2e_xtc scaling; no CLI leg, no CPU baseline)
O=gpurun_out/r2/xtc_n4; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-cli > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"; tail -3 $O/bench_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2/xtc_n4/bench_n4.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "trr", d["e2e_trr"]["value"], "xtc", (d.get("e2e_xtc") or {}).get("value"), "allreduce", d["allreduce_check"]["ok"] if d.get("allreduce_check") else None)
PY
