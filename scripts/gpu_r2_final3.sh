#!/bin/bash
# round 2, last one-GPU evidence run (K1x's index pass on its own stream): the whole GPU test-suite, smoke, the default bench line
O=gpurun_out/r2/final3; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2/final3/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "trr", d["e2e_trr"]["value"], "xtc", (d.get("e2e_xtc") or {}).get("value"), "roofline", d["roofline"]["frac"])
print("cli", (d.get("e2e_cli") or {}).get("value"), (d.get("e2e_cli") or {}).get("in_loop_frames_per_s"))
PY
