#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/attn_tests2.log
rm -f gpurun_out/attn_ab2.log
cp ab_tmp/libdtb200_prev.so /tmp/prev.so
for rep in 1 2; do
for lib in /tmp/prev.so ""; do
  for cfg in "--batch 8 --seq 512 --dropout 0" "--batch 1 --seq 512 --dropout 0"; do
    DTB200_LIB=$lib timeout 300 python scripts/step_bench.py --model gpt2 $cfg --steps 30 2>&1 | tail -1 | sed "s|^|lib=${lib:-new} |" | tee -a gpurun_out/attn_ab2.log
  done
done
done
