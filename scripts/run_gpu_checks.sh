set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 400 python bench.py --gpus 1 --steps 30 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n1_dropout.jsonl | cut -c1-1800
timeout 600 python bench.py --impl reference --gpus 1 --steps 8 --warmup 3 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_reference_n1_b512.jsonl | cut -c1-1200
