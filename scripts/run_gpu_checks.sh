set -x
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine.py tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -4
DTB200_GEMM_CLUSTER=1 timeout 600 python scripts/gemm_check.py --only perf 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print('CL1', r['name'], r.get('ok'), round(r.get('tflops',0)), round(r.get('cublas_tflops',0)))
    except Exception: print(l[:300])"
timeout 900 python scripts/gemm_check.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print('CL2', r['name'], r.get('ok'), round(r.get('tflops',0)), round(r.get('cublas_tflops',0)), r.get('error','')[:200])
    except Exception: print(l[:300])"
timeout 300 python scripts/step_bench.py --batch 256 2>&1 | tail -1
timeout 600 python scripts/step_bench.py --model llama-3.2-1b --batch 8 --seq 512 --steps 10 --warmup 3 --lm-chunk 4096 2>&1 | tail -2
for m in sys nc weak; do DTB200_PEER_LD=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/bandwidth_sweep.py --sizes-mb 1024 2>&1 | grep SWEEP | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l[6:]); print('$m', {k: round(v,3) if isinstance(v,float) else v for k,v in r.items() if 'ms_' in k or 'gbs' in k})"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/validator_bench.py --model gpt2-medium 2>&1 | grep -E "VALBENCH|Error|error" | cut -c1-1500
