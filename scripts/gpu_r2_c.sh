#!/usr/bin/env bash
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "adamw or meta or padding or hf" 2>&1 | tail -6
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $R --master-port 29561 scripts/validator_bench.py --model gpt2-medium 2>&1 | grep -E "VALBENCH|rror|Trace|File" | cut -c1-3000
timeout 600 $R --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_n${N}_c.jsonl | cut -c1-3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b512.csv python scripts/step_bench.py --batch 512 --steps 1 --warmup 1 --no-graph > gpurun_out/step_ncu.log 2>&1
python scripts/kernel_shares.py gpurun_out/launches_b512.csv > gpurun_out/kernel_shares_r2_b512.json 2>&1; head -c 3000 gpurun_out/kernel_shares_r2_b512.json
