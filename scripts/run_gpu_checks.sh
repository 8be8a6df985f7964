set -x
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $R --master-port 29531 scripts/bcast_gemm_bench.py 2>&1 | grep -E "BCASTGEMM|rror" | cut -c1-1500
timeout 600 $R --master-port 29532 bench.py --gpus 2 --steps 12 --warmup 3 --model llama-3.2-1b --batch-size 8 --seq-len 512 --delta-dtype fp8 --fp8-forward --no-e2e 2>&1 | grep -E '^\{|rror' | cut -c1-1500
