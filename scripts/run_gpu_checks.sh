set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sm100_gemm_kernel" --launch-skip 152 -c 3 -f -o gpurun_out/gemm_fwd3 python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sm100_gemm_kernel" --launch-skip 198 -c 2 -f -o gpurun_out/gemm_bwd2 python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_new.csv python scripts/step_bench.py --batch 256 --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
