#!/usr/bin/env bash
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $R --master-port 29581 scripts/meta_check.py --model gpt2 --val-batch 8 --val-seq 512 --steps 12 --skip-collective 2>&1 | grep -E "META_CHECK|rror|Trace|File|line " | cut -c1-4500
timeout 600 $R --master-port 29582 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_n${N}_d.jsonl | cut -c1-3500
timeout 900 python -m pytest tests -x -q -m gpu -k "multigpu or adamw or meta" 2>&1 | tail -8
