set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_small|attn_bwd_small|ce_fwd_bwd_smem|norm_fwd_fast|norm_bwd_fast|colsum|adamw" --launch-skip 70 -c 8 -f -o /tmp/misc3 python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_m.log 2>&1
python scripts/ncu_summary.py /tmp/misc3.ncu-rep > gpurun_out/ncu_misc_v3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sm100_gemm_kernel" --launch-skip 151 -c 4 -f -o /tmp/gemm3 python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_g.log 2>&1
python scripts/ncu_summary.py /tmp/gemm3.ncu-rep > gpurun_out/ncu_gemm_v3.json
ls -la gpurun_out/ncu_*_v3.json
