#!/usr/bin/env bash
# round-2 N-GPU campaign (run once at N=8): every multi-GPU number quoted in RESULTS.md / profiles/
set -x
N=${N:-8}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $R --master-port 29701 scripts/vmm_check.py 2>&1 | grep -E "VMM_CHECK|rror" | cut -c1-600
timeout 500 $R --master-port 29702 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{|rror|Trace|File|line ' | tee gpurun_out/bench_n$N.jsonl | cut -c1-4000
timeout 400 $R --master-port 29703 bench.py --gpus $N --steps 6 --warmup 3 --no-e2e --meta-mode dp 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_n${N}_metadp.jsonl | cut -c1-3000
timeout 400 $R --master-port 29704 scripts/meta_check.py --model gpt2 --val-batch 8 --val-seq 512 --steps 12 2>&1 | grep -E "META_CHECK|rror|Trace" | cut -c1-6000
timeout 400 $R --master-port 29705 scripts/meta_check.py --model gpt2 --val-batch 96 --val-seq 512 --steps 8 --skip-collective --out gpurun_out/meta_check_n${N}_gpt2_b96.json 2>&1 | grep -E "META_CHECK|rror|Trace" | cut -c1-4000
timeout 500 $R --master-port 29706 scripts/validator_bench.py --model gpt2-medium 2>&1 | grep -E "VALBENCH|rror|Trace" | cut -c1-3000
timeout 600 $R --master-port 29707 scripts/bandwidth_sweep.py --sizes-mb 1,64,1024,4096 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-2200
timeout 300 $R --master-port 29708 scripts/bandwidth_sweep.py --sizes-mb 64,1024 --dtype bf16 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-2200
timeout 300 $R --master-port 29709 scripts/bandwidth_sweep.py --sizes-mb 64,1024 --dtype fp8 2>&1 | grep -E "SWEEP|rror|Trace" | cut -c1-2200
timeout 500 $R --master-port 29710 bench.py --gpus $N --impl nccl --steps 6 --warmup 3 --no-e2e --val-texts 16 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_nccl_n$N.jsonl | cut -c1-3000
timeout 500 $R --master-port 29711 bench.py --gpus $N --model llama-3.2-1b --batch-size 8 --seq-len 512 --fp8-forward --delta-dtype fp8 --steps 10 --warmup 3 --no-full-round --local-steps 5 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_llama_fp8_n$N.jsonl | cut -c1-2500
timeout 500 $R --master-port 29712 bench.py --gpus $N --model llama-3.2-1b --batch-size 8 --seq-len 512 --delta-dtype bf16 --steps 10 --warmup 3 --no-full-round --local-steps 5 2>&1 | grep -E '^\{|rror|Trace' | tee gpurun_out/bench_llama_bf16_n$N.jsonl | cut -c1-2500
