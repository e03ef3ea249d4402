#!/bin/bash
# K3 tuning: ring depth / frames per stage (rebuilds the library with -D overrides on the GPU box)
run() { python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('  ', d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['covariance_max_scaled_error_vs_numpy'])"; }
echo "default (K 16, 5 stages)"; run
for v in ${VARIANTS:-"-DCGC_SYRK_STAGES=6" "-DCGC_SYRK_STAGES=4" "-DCGC_SYRK_K=32 -DCGC_SYRK_STAGES=3" "-DCGC_SYRK_K=8 -DCGC_SYRK_STAGES=8"}; do
  echo "variant $v"
  python -c "
from cgromcorrfmo_b200 import build
build.NVCC_FLAGS += '$v'.split()
build.build(force=True)" > /dev/null 2>&1
  run
done
python -c "
from cgromcorrfmo_b200 import build
build.build(force=True)" > /dev/null 2>&1
