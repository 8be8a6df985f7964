set -x
timeout 600 python scripts/gemm_check.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print('2SM', r['name'], r.get('ok'), round(r.get('tflops',0)), round(r.get('cublas_tflops',0)), r.get('error','')[:300])
    except Exception: print(l[:300])"
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/step_bench.py --batch 256 2>&1 | tail -1
DTB200_GEMM_CLUSTER=1 timeout 300 python scripts/step_bench.py --batch 256 2>&1 | tail -1
