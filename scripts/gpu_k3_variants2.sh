#!/bin/bash
run() { python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline $1 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('  ', d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['covariance_max_scaled_error_vs_numpy'])"; }
for v in "-DCGC_SYRK_K=32 -DCGC_SYRK_STAGES=2" "-DCGC_SYRK_K=24 -DCGC_SYRK_STAGES=4" "-DCGC_SYRK_K=48 -DCGC_SYRK_STAGES=2" "-DCGC_SYRK_K=32 -DCGC_SYRK_STAGES=3"; do
  echo "variant $v"
  python -c "
from cgromcorrfmo_b200 import build
build.NVCC_FLAGS += '$v'.split()
build.build(force=True)" > /dev/null 2>&1
  run
  if [ "$v" == "-DCGC_SYRK_K=32 -DCGC_SYRK_STAGES=3" ]; then echo "  same, 256 frames per launch"; run "--frames-per-step 256"; fi
done
