set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "fp8 or gemm" 2>&1 | tail -3
timeout 300 $R --master-port 29521 scripts/peer_check.py 2>&1 | grep -E "PEER_CHECK|rror" | cut -c1-1800
timeout 400 $R --master-port 29524 scripts/bandwidth_sweep.py --sizes-mb 16,1024 2>&1 | grep -E "SWEEP|rror" | cut -c1-1500
timeout 400 $R --master-port 29522 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e 2>&1 | grep -E '^\{|rror' | cut -c1-700
timeout 400 $R --master-port 29523 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --impl nccl 2>&1 | grep -E '^\{|rror' | cut -c1-700
timeout 600 $R --master-port 29526 bench.py --gpus 2 --steps 6 --warmup 3 --impl reference 2>&1 | grep -E '^\{|rror' | cut -c1-900
