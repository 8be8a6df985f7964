#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine.py tests/test_hf_parity.py -x -q -m gpu 2>&1 | tail -4
python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1
DTB200_GEMM_NO_EPI_SPEC=1 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1
python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:norm_bwd_fast --launch-skip 4 -c 3 python scripts/step_bench.py --batch 512 --steps 1 --warmup 0 --no-graph 2>&1 | grep norm_bwd | cut -c1-120,300-420
