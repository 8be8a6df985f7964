set -x
N=${N:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 500 $R --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n$N.jsonl | cut -c1-1800
timeout 300 $R --master-port 29542 bench.py --gpus $N --steps 40 --warmup 5 --impl nccl --no-e2e 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_nccl_n$N.jsonl | cut -c1-600
timeout 300 $R --master-port 29543 scripts/peer_check.py 2>&1 | grep -E "PEER_CHECK|rror" | cut -c1-1500
timeout 400 $R --master-port 29544 scripts/bandwidth_sweep.py 2>&1 | grep -E "SWEEP|rror" | cut -c1-2500
timeout 300 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -3
