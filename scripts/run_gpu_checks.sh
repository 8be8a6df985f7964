#!/usr/bin/env bash
# The GPU-side check list used during development (run on a B200 box from the repo root):
#   bash scripts/run_gpu_checks.sh            # 1 GPU: kernel + engine tests, smoke, step time, headline bench
#   N=2 bash scripts/run_gpu_checks.sh        # additionally the multi-GPU planes (peer windows, NVLS) and the N-GPU bench
set -x
N=${N:-1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1
if [ "$N" -gt 1 ]; then
  R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 300 $R --master-port 29541 scripts/peer_check.py 2>&1 | grep -E "PEER_CHECK|rror" | cut -c1-1500
  timeout 300 $R --master-port 29542 scripts/nvls_check.py 2>&1 | grep -E "NVLS_CHECK|rror" | cut -c1-1500
  timeout 500 $R --master-port 29543 bench.py --gpus $N --steps 40 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n$N.jsonl | cut -c1-2000
else
  timeout 400 python bench.py --gpus 1 --steps 40 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n1.jsonl | cut -c1-2000
fi
