#!/usr/bin/env bash
# ncu --set full of one SMALL GEMM of the learned-mixer step (out-projection, M = 4096, N = K = 768): where do the ~10 us of fixed cost go?
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel --launch-skip 1 -c 1 -o /tmp/gemm_small -f \
   python scripts/step_bench.py --batch 8 --seq 512 --steps 1 --warmup 0 --no-graph --dropout 0 > gpurun_out/ncu_gemm_small.log 2>&1
python scripts/ncu_summary.py /tmp/gemm_small.ncu-rep > gpurun_out/ncu_r2_gemm_small_oproj.json 2>&1
ncu -i /tmp/gemm_small.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_r2_gemm_small_oproj_source.csv 2>/dev/null
ncu -i /tmp/gemm_small.ncu-rep --page details --csv > gpurun_out/ncu_r2_gemm_small_oproj_details.csv 2>/dev/null
head -c 1500 gpurun_out/ncu_r2_gemm_small_oproj.json
