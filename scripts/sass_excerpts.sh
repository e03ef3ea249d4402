#!/bin/bash
# SASS evidence for profiles/: per-kernel opcode histograms of the shipped library plus excerpts of K2's pair loop and K3's
# main loop (cuobjdump -sass; run after python -m cgromcorrfmo_b200.build).
LIB=cgromcorrfmo_b200/libcgromcorr.so
OUT=${1:-profiles/r02_sass_excerpts.txt}
name() { cuobjdump -sass $LIB 2>/dev/null | grep "Function :" | grep -E "$1" | head -1 | sed 's/.*Function : //'; }
body() { cuobjdump -sass -fun "$1" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'; }
K2=$(name 'cdc_kernelILi2ELi42ELi0'); K3=$(name 'cov_syrk_kernel'); K1=$(name 'decode_kernelILi1')
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
  echo "# resource usage (registers, shared memory) of the hot-path kernels"
  cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 -E "Function .*(cdc_kernelILi2ELi42ELi0|decode_kernelILi1|cov_syrk_kernel|cov_centre_kernel|finish_kernel|format_rows_kernel)" | grep -E "Function|REG"
  for k in "$K2" "$K3" "$K1"; do
    echo; echo "## opcode histogram: $k"
    body "$k" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/^([A-Z0-9_]+(\.(RSQ|E|ADD|F64|64|8x8x4|3D|128))*).*/\1/' | sort | uniq -c | sort -rn | head -26
  done
  echo; echo "## K2 cdc_kernel<2,42,0>: 56 instructions from the middle of the unrolled pair loop (FADD2/FMUL2/FFMA2 packed FP32, MUFU.RSQ, broadcast LDS.128)"
  L=$(body "$K2" | grep -n "MUFU.RSQ" | sed -n '120p' | cut -d: -f1)
  body "$K2" | sed -n "${L},$((L+56))p"
  echo; echo "## K2: site blocks by bulk TMA (UBLKCP) on an mbarrier ring (SYNCS), bins by fire-and-forget FP64 reductions (REDG.E.ADD.F64) after grid rounding (DADD)"
  body "$K2" | grep -E "REDG|UBLKCP|SYNCS" | head -12
  echo; echo "## K3 cov_syrk_kernel: tensor-map TMA copies (UTMALDG.3D) and one k-step of the main loop (8 LDS.64 feeding 16 DMMA.8x8x4)"
  body "$K3" | grep -E "UTMALDG" | head -4
  L=$(body "$K3" | grep -n "DMMA" | sed -n '17p' | cut -d: -f1)
  body "$K3" | sed -n "$((L-12)),$((L+22))p"
  echo; echo "## K3 write-back: RED.E.ADD.F64"
  body "$K3" | grep -E "REDG" | head -3
  echo; echo "## K1 decode_kernel<1>: bulk TMA copy + mbarrier, PRMT realignment"
  body "$K1" | grep -E "UBLKCP|SYNCS|PRMT" | head -8
} > $OUT
wc -l $OUT
