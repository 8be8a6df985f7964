#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for B in 1 8; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_t512_b$B.csv python scripts/step_bench.py --batch $B --seq 512 --steps 1 --warmup 1 --no-graph --dropout 0 > /dev/null 2>&1
python scripts/kernel_shares.py gpurun_out/launches_t512_b$B.csv > gpurun_out/kernel_shares_t512_b$B.json
python - <<PY
import json
d=json.load(open("gpurun_out/kernel_shares_t512_b$B.json"))
print("B=$B total_us(2 steps)", d["total_us"])
for k in d["kernels"][:12]: print("  ", k["kernel"][:70], k["launches"], round(k["total_us"]), round(k["share"],3))
PY
python scripts/step_bench.py --batch $B --seq 512 --steps 20 --dropout 0 2>&1 | tail -1 | cut -c1-200
done
