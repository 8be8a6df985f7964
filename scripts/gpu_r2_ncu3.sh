#!/usr/bin/env bash
# ncu --set full of the tiled (T = 512) attention kernels of the learned-mixer step
set -x
mkdir -p gpurun_out
for K in attn_fwd_kernel attn_bwd_kv_kernel attn_bwd_q_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 5 -c 1 -o /tmp/$K -f \
     python scripts/step_bench.py --batch 8 --seq 512 --steps 1 --warmup 0 --no-graph > gpurun_out/ncu_$K.log 2>&1
  python scripts/ncu_summary.py /tmp/$K.ncu-rep > gpurun_out/ncu_r2_$K.json 2>&1
  ncu -i /tmp/$K.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_r2_${K}_source.csv 2>/dev/null
done
