#!/usr/bin/env bash
# round-2 GPU check B (N GPUs): all gpu tests incl. multi-GPU, distributed meta-learner check + timing, N-GPU bench
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_n$N.txt
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $R --master-port 29551 scripts/meta_check.py --model gpt2 --val-batch 8 --val-seq 512 --steps 12 2>&1 | grep -E "META_CHECK|rror|Trace" | cut -c1-4000
timeout 400 $R --master-port 29552 scripts/meta_check.py --model gpt2 --val-batch 96 --val-seq 512 --steps 8 --skip-collective --out gpurun_out/meta_check_n${N}_gpt2_b96.json 2>&1 | grep -E "META_CHECK|rror|Trace" | cut -c1-3000
timeout 600 $R --master-port 29553 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_n$N.jsonl | cut -c1-3500
