set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n2_final.jsonl | cut -c1-1700
