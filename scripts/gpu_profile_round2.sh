set -x
python scripts/prof_k5_k8.py || exit 1
for k in format_rows_kernel parse_kernel parse_count_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python scripts/prof_k5_k8.py > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:jacobi_block -s 150 -c 1 -f -o gpurun_out/prof_k6_block_dmma python scripts/dev_eigh_check.py > gpurun_out/ncu_k6_dmma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_kernel -c 1 -f -o gpurun_out/prof_k4_project python scripts/dev_project_check.py 1024 > gpurun_out/ncu_k4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
