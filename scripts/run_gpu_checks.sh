set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step2.csv python scripts/step_bench.py --batch 256 --no-graph --steps 1 --warmup 1 > gpurun_out/step_ncu2.log 2>&1
python scripts/kernel_shares.py gpurun_out/launches_step2.csv > gpurun_out/kernel_shares_v2.json
ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel -s 30 -c 6 -o gpurun_out/prof_gemm python scripts/step_bench.py --batch 256 --no-graph --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -o gpurun_out/prof_attn python scripts/step_bench.py --batch 256 --no-graph --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/
