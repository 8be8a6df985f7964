set -x
timeout 400 python bench.py --gpus 1 --steps 12 --warmup 3 2>&1 | grep -E '^\{|rror|Traceback' | cut -c1-2200
