#!/usr/bin/env bash
# one ncu --set full capture per hot kernel (1 GPU), summarised on the box; reports stay in /tmp unless small
set -x
mkdir -p gpurun_out
for K in attn_bwd_small_kernel norm_bwd_fast_kernel ce_fwd_bwd_smem_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 -c 1 -o /tmp/$K -f \
     python scripts/step_bench.py --batch 512 --steps 1 --warmup 0 --no-graph > gpurun_out/ncu_$K.log 2>&1
  python scripts/ncu_summary.py /tmp/$K.ncu-rep > gpurun_out/ncu_r2_$K.json 2>&1
  ncu -i /tmp/$K.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_r2_${K}_source.csv 2>/dev/null
  ncu -i /tmp/$K.ncu-rep --page details --csv > gpurun_out/ncu_r2_${K}_details.csv 2>/dev/null
  ls -la /tmp/$K.ncu-rep
done
