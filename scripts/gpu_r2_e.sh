#!/usr/bin/env bash
set -x
N=${N:-2}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $R --master-port 29591 bench.py --gpus $N --model llama-3.2-1b --batch-size 8 --seq-len 512 --fp8-forward --delta-dtype fp8 --steps 8 --warmup 3 --no-full-round --local-steps 4 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_llama_fp8_n$N.jsonl | cut -c1-2500
timeout 600 $R --master-port 29592 bench.py --gpus $N --model llama-3.2-1b --batch-size 8 --seq-len 512 --delta-dtype bf16 --steps 8 --warmup 3 --no-full-round --local-steps 4 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_llama_bf16_n$N.jsonl | cut -c1-2500
timeout 600 $R --master-port 29593 bench.py --gpus $N --impl nccl --steps 6 --warmup 3 --no-e2e --val-texts 16 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_nccl_n$N.jsonl | cut -c1-2500
timeout 600 $R --master-port 29594 bench.py --gpus $N --impl torch-bf16 --steps 6 --warmup 3 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_torch_bf16_n$N.jsonl | cut -c1-1500
timeout 600 python bench.py --impl reference --gpus 1 --steps 4 --warmup 3 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_reference_n1.jsonl | cut -c1-1500
