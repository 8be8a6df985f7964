mkdir -p gpurun_out; rm -f gpurun_out/ab3.log
for rep in 1 2; do
python scripts/step_bench.py --batch 512 --seq 64 --steps 30 --lm-chunk 16384 2>&1 | tail -1 | sed 's/^/chunk16384 /' | tee -a gpurun_out/ab3.log
python scripts/step_bench.py --batch 512 --seq 64 --steps 30 --lm-chunk 32768 2>&1 | tail -1 | sed 's/^/chunk32768 /' | tee -a gpurun_out/ab3.log
python scripts/step_bench.py --batch 512 --seq 64 --steps 30 --lm-chunk 8192 2>&1 | tail -1 | sed 's/^/chunk8192 /' | tee -a gpurun_out/ab3.log
python scripts/step_bench.py --batch 8 --seq 512 --steps 30 --dropout 0 2>&1 | tail -1 | sed 's/^/t512 /' | tee -a gpurun_out/ab3.log
DTB200_PDL_ALL=1 python scripts/step_bench.py --batch 8 --seq 512 --steps 30 --dropout 0 2>&1 | tail -1 | sed 's/^/t512_pdlall /' | tee -a gpurun_out/ab3.log
done
