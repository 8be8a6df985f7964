set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 40 --warmup 5 2>&1 | grep -E '^\{|rror' | tee gpurun_out/bench_n1_final.jsonl | cut -c1-1800
