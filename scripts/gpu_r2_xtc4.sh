#!/bin/bash
# round 2: ncu --set full of the two .xtc kernels (one launch each), for the stall picture of the index walk
O=gpurun_out/r2/xtc4; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xtc_index_kernel -s 3 -c 1 -f -o $O/xtc_index python scripts/dev_xtc_check.py 256 0 256 > $O/ncu_index.log 2>&1; echo "ncu index rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xtc_decode_kernel -s 4 -c 1 -f -o $O/xtc_decode python scripts/dev_xtc_check.py 256 0 256 > $O/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
ls -la $O
