set -x
P='import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(sys.argv[1], r["name"], r.get("ok"), round(r.get("tflops",0)), round(r.get("cublas_tflops",0)), r.get("error","")[:200])
    except Exception: print(l[:200])'
timeout 600 python scripts/gemm_check.py 2>&1 | python -c "$P" NEW
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
(cd gpurun_tmp/wt_2sm && timeout 300 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-130)
timeout 300 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-130
timeout 300 python scripts/step_bench.py --batch 512 --steps 20 2>&1 | tail -1 | cut -c1-130
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step5.csv python scripts/step_bench.py --batch 256 --no-graph --steps 1 --warmup 1 > gpurun_out/step_ncu5.log 2>&1
python scripts/kernel_shares.py gpurun_out/launches_step5.csv > gpurun_out/kernel_shares_v5.json
