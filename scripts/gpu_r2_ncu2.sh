#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for SK in 54 2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel --launch-skip $SK -c 1 -o /tmp/gemm_$SK -f \
     python scripts/step_bench.py --batch 512 --steps 1 --warmup 0 --no-graph > gpurun_out/ncu_gemm_$SK.log 2>&1
  python scripts/ncu_summary.py /tmp/gemm_$SK.ncu-rep > gpurun_out/ncu_r2_gemm_skip$SK.json 2>&1
  ncu -i /tmp/gemm_$SK.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_r2_gemm_skip${SK}_source.csv 2>/dev/null
  ncu -i /tmp/gemm_$SK.ncu-rep --page details --csv > gpurun_out/ncu_r2_gemm_skip${SK}_details.csv 2>/dev/null
done
