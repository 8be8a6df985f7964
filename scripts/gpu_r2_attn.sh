#!/bin/bash
# attention softmax rewrite: numerics + A/B against the previous library (DTB200_LIB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_engine.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/attn_tests.log
rm -f gpurun_out/attn_ab.log
for rep in 1 2; do
for lib in ab_tmp/libdtb200_old.so ""; do
  for cfg in "--batch 8 --seq 512 --dropout 0" "--batch 512 --seq 64"; do
    DTB200_LIB=$lib timeout 300 python scripts/step_bench.py --model gpt2 $cfg --steps 30 2>&1 | tail -1 | sed "s|^|lib=${lib:-new} |" | tee -a gpurun_out/attn_ab.log
  done
done
done
