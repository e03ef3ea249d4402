#!/bin/bash
# round 2: K1x's index pass on a stream of its own (beside the previous batch's pair kernel) against both passes in the batch's turn
O=gpurun_out/r2/xtc7; mkdir -p $O
timeout 400 python -m pytest tests/test_xtc.py tests/test_trr.py -m gpu -x -q > $O/pytest_xtc.log 2>&1; echo "pytest xtc rc=$?"; tail -3 $O/pytest_xtc.log
for ov in 1 0; do
  CGC_XTC_OVERLAP=$ov timeout 200 python scripts/dev_xtc_check.py 512 0 64,128 >> $O/dev_xtc_check.log 2>&1; echo "check ov=$ov rc=$?"
  CGC_XTC_OVERLAP=$ov timeout 300 python scripts/dev_xtc_check.py 512 0 64,128 water >> $O/dev_xtc_check_water.log 2>&1; echo "water ov=$ov rc=$?"
done
grep overlap $O/dev_xtc_check.log $O/dev_xtc_check_water.log | cut -c1-230
