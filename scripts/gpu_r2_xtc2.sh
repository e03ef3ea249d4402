#!/bin/bash
# round 2: the .xtc path after the index pass moved its stream into a shared-memory ring: tests, kernel times, per-kernel split (ncu launch list)
O=gpurun_out/r2/xtc6; mkdir -p $O
timeout 400 python -m pytest tests/test_xtc.py -m gpu -x -q > $O/pytest_xtc.log 2>&1; echo "pytest xtc rc=$?"; tail -4 $O/pytest_xtc.log
timeout 300 python scripts/dev_xtc_check.py 256 0 64,128,256 > $O/dev_xtc_check.log 2>&1; echo "dev check rc=$?"; tail -5 $O/dev_xtc_check.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:xtc -c 40 --csv --log-file $O/xtc_launches.csv python scripts/dev_xtc_check.py 256 0 256 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/xtc6/xtc_launches.csv")) if len(r) > 5 and r[0].isdigit()]
t = collections.defaultdict(list)
for r in rows:
    t[r[4].split("(")[0]].append(float(r[-1].replace(",", "")))
for k, v in t.items():
    print(k, "launches", len(v), "min/median/max us", min(v), sorted(v)[len(v)//2], max(v), "(unit as ncu reports)")
PY
