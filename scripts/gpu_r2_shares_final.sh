#!/usr/bin/env bash
# per-kernel time shares of the final miner step (B = 512 x 64) and of the learned-mixer batch shape (8 x 512)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b512_final.csv python scripts/step_bench.py --batch 512 --steps 1 --warmup 1 --no-graph > /dev/null 2>&1
python scripts/kernel_shares.py gpurun_out/launches_b512_final.csv > gpurun_out/kernel_shares_r2_b512_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_t512_b8_final.csv python scripts/step_bench.py --batch 8 --seq 512 --steps 1 --warmup 1 --no-graph --dropout 0 > /dev/null 2>&1
python scripts/kernel_shares.py gpurun_out/launches_t512_b8_final.csv > gpurun_out/kernel_shares_r2_t512_b8_final.json
python - <<PY
import json
for f in ("kernel_shares_r2_b512_final","kernel_shares_r2_t512_b8_final"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "total_us(2 steps)", d["total_us"])
    for k in d["kernels"][:10]: print("  ", k["kernel"][:60], k["launches"], round(k["total_us"]), round(k["share"],3))
PY
