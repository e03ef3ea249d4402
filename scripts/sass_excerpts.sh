#!/bin/bash
# SASS evidence for profiles/: per-kernel opcode histograms of the shipped library plus excerpts of K2's pair loop and K3's
# main loop (cuobjdump -sass; run after python -m cgromcorrfmo_b200.build).
LIB=cgromcorrfmo_b200/libcgromcorr.so
OUT=${1:-profiles/r02_sass_excerpts.txt}
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
  echo "# resource usage (registers, shared memory) of the hot-path kernels"
  cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 -E "Function .*(cdc_kernelILi2ELi42ELi0|decode_kernelILi1|cov_syrk_kernel|cov_sum_kernel|finish_kernel|format_rows_kernel)" | grep -E "Function|REG"
  for k in '_ZN3cgc10cdc_kernelILi2ELi42ELi0EEEvPKfPK6float4PKdPKiPxPiiixiiiii' '_ZN3cgc15cov_syrk_kernelE14CUtensorMap_stiPdiiii'; do
    echo; echo "## opcode histogram: $k"
    cuobjdump -sass -fun "$k" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/^([A-Z0-9_]+(\.(RSQ|E|ADD|F64|64|884|RCP))*).*/\1/' | sort | uniq -c | sort -rn | head -28
  done
  echo; echo "## K2 cdc_kernel<2,42,0>: 60 instructions from the middle of the unrolled pair loop (FADD2/FMUL2/FFMA2 packed FP32, MUFU.RSQ, broadcast LDS.128)"
  cuobjdump -sass -fun '_ZN3cgc10cdc_kernelILi2ELi42ELi0EEEvPKfPK6float4PKdPKiPxPiiixiiiii' $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | grep -n "MUFU.RSQ" | sed -n '120p' | cut -d: -f1 > /tmp/k2_line
  L=$(cat /tmp/k2_line)
  cuobjdump -sass -fun '_ZN3cgc10cdc_kernelILi2ELi42ELi0EEEvPKfPK6float4PKdPKiPxPiiixiiiii' $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -n "$((L)),$((L+60))p" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
  echo; echo "## K2 epilogue: fixed-point bins (DFMA magic-number conversion, REDG.E.ADD.64 integer reductions; no F2I)"
  cuobjdump -sass -fun '_ZN3cgc10cdc_kernelILi2ELi42ELi0EEEvPKfPK6float4PKdPKiPxPiiixiiiii' $LIB 2>/dev/null | grep -E "REDG|UBLKCP|SYNCS" | head -12 | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
  echo; echo "## K3 cov_syrk_kernel: TMA issue (UTMALDG) and one k-step of the main loop (12 LDS.64 feeding 32 DMMA.884)"
  cuobjdump -sass -fun '_ZN3cgc15cov_syrk_kernelE14CUtensorMap_stiPdiiii' $LIB 2>/dev/null | grep -E "UTMALDG" | head -4 | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
  L=$(cuobjdump -sass -fun '_ZN3cgc15cov_syrk_kernelE14CUtensorMap_stiPdiiii' $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | grep -n "DMMA" | sed -n '33p' | cut -d: -f1)
  cuobjdump -sass -fun '_ZN3cgc15cov_syrk_kernelE14CUtensorMap_stiPdiiii' $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -n "$((L-14)),$((L+36))p" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
  echo; echo "## K1 decode_kernel<1>: bulk TMA copy + mbarrier"
  cuobjdump -sass $LIB 2>/dev/null | awk '/Function : .*decode_kernelILi1/{f=1} /Function : /{if(!/decode_kernelILi1/)f=0} f' | grep -E "UBLKCP|SYNCS|PRMT" | head -8 | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
} > $OUT
wc -l $OUT
