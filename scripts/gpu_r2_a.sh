#!/usr/bin/env bash
# round-2 GPU check A (1 GPU): all gpu tests, smoke, headline bench incl. the directly measured whole round, torch-bf16 arm
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace|File' | tee gpurun_out/bench_n1.jsonl | cut -c1-3500
timeout 400 python bench.py --impl torch-bf16 --gpus 1 --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_torch_bf16_n1.jsonl | cut -c1-1500
timeout 300 python scripts/meta_check.py --model gpt2 --val-batch 8 --val-seq 512 --steps 10 2>&1 | grep -E "META_CHECK|rror|Trace" | cut -c1-2500
