#!/bin/bash
# round 2: first run of the .xtc path on a B200: its GPU tests, kernel times, the bench line with the e2e_xtc leg
O=gpurun_out/r2/xtc1; mkdir -p $O
timeout 400 python -m pytest tests/test_xtc.py -m gpu -x -q > $O/pytest_xtc.log 2>&1; echo "pytest xtc rc=$?"; tail -15 $O/pytest_xtc.log
timeout 300 python scripts/dev_xtc_check.py 256 > $O/dev_xtc_check.log 2>&1; echo "dev check rc=$?"; cat $O/dev_xtc_check.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cli > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2/xtc1/bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value",)}, "e2e", d["e2e"]["value"], "trr", d["e2e_trr"]["value"])
print("xtc", json.dumps(d.get("e2e_xtc"))[:1500])
PY
