#!/usr/bin/env bash
set -x
N=${N:-4}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $R --master-port 29721 scripts/bandwidth_sweep.py --sizes-mb 1,64,1024,4096 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-300
timeout 300 $R --master-port 29722 scripts/bandwidth_sweep.py --sizes-mb 64,1024 --dtype bf16 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-300
timeout 300 $R --master-port 29723 scripts/bandwidth_sweep.py --sizes-mb 64,1024 --dtype fp8 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-300
timeout 500 $R --master-port 29724 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_n$N.jsonl | cut -c1-3000
