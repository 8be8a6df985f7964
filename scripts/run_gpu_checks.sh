set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 scripts/nvls_check.py 2>&1 | grep -E "NVLS_CHECK|rror|Traceback|File " | cut -c1-1200 | tail -12
