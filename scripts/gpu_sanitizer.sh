#!/bin/bash
# compute-sanitizer over the hot-path kernels: smoke() + the small parity tests, three tools.
# Summaries go to gpurun_out/r2/sanitizer_<tool>.log; copy them to profiles/ after reading.
mkdir -p gpurun_out/r2
SEL="dE_against_oracle or test_covariance or random_shapes or rows_bit_identical or write_time_one_call or other_dq_atom_counts or decode_bit_exact"
for tool in memcheck racecheck synccheck; do
  { echo "== compute-sanitizer --tool $tool : smoke()"; 
    timeout 600 compute-sanitizer --tool $tool --print-limit 30 python __graft_entry__.py smoke 2>&1 | tail -25
    echo "== compute-sanitizer --tool $tool : pytest -k '$SEL'"
    timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python -m pytest tests/test_gpu_parity.py tests/test_engine.py -m gpu -q -k "$SEL" 2>&1 | tail -40
  } > gpurun_out/r2/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2/sanitizer_$tool.log | tail -6
done
