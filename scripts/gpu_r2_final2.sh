#!/bin/bash
# round 2, second final one-GPU evidence run (after the .xtc path was added): the whole GPU test-suite, smoke, the default bench line,
# ncu --set full of the two K1x kernels
O=gpurun_out/r2/final2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"; tail -3 $O/bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xtc_index_kernel -s 2 -c 1 -f -o $O/xtc_index python scripts/dev_xtc_check.py 256 0 256 > $O/ncu_index.log 2>&1; echo "ncu index rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xtc_decode_kernel -s 2 -c 1 -f -o $O/xtc_decode python scripts/dev_xtc_check.py 256 0 256 > $O/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 300 python scripts/dev_xtc_check.py 256 0 64,256 water > $O/dev_xtc_check_water.log 2>&1; echo "water check rc=$?"; tail -3 $O/dev_xtc_check_water.log
timeout 200 python scripts/dev_xtc_check.py 256 0 64,256 > $O/dev_xtc_check.log 2>&1; echo "check rc=$?"; tail -3 $O/dev_xtc_check.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2/final2/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "trr", d["e2e_trr"]["value"], "xtc", (d.get("e2e_xtc") or {}).get("value"), "roofline", d["roofline"]["frac"])
print("cli", (d.get("e2e_cli") or {}).get("value"), (d.get("e2e_cli") or {}).get("in_loop_frames_per_s"))
PY
