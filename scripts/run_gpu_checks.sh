set -x
P='import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(sys.argv[1], r["name"], r.get("ok"), round(r.get("tflops",0)), round(r.get("cublas_tflops",0)), r.get("error","")[:200])
    except Exception: print(l[:200])'
(cd gpurun_tmp/wt_2sm && timeout 300 python scripts/gemm_check.py --only perf 2>&1 | python -c "$P" OLD)
timeout 600 python scripts/gemm_check.py 2>&1 | python -c "$P" NEW
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine.py -m gpu -x -q 2>&1 | tail -2
(cd gpurun_tmp/wt_2sm && timeout 300 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-130)
timeout 300 python scripts/step_bench.py --batch 256 --steps 30 2>&1 | tail -1 | cut -c1-130
