#!/bin/bash
# fp8 dgrad: numerics tests + Llama / GPT-2 step A/B (bf16 | fp8 fwd | fp8 fwd + dgrad)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine.py tests/test_ops_gpu.py -q -m gpu -x -k "fp8" 2>&1 | tail -15 > gpurun_out/fp8bwd_tests.log
cat gpurun_out/fp8bwd_tests.log
for f in "" "--fp8" "--fp8 --fp8-bwd"; do
  timeout 300 python scripts/step_bench.py --model llama-3.2-1b --batch 8 --seq 512 $f 2>&1 | tail -1 | tee -a gpurun_out/fp8bwd_step.log
done
for f in "" "--fp8" "--fp8 --fp8-bwd"; do
  timeout 300 python scripts/step_bench.py --model gpt2 --batch 512 --seq 64 $f 2>&1 | tail -1 | tee -a gpurun_out/fp8bwd_step.log
done
