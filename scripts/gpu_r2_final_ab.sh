#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/final_ab.log
for rep in 1 2 3; do
  for v in 1 0; do
  DTB200_ATEN_SMALL=$v python scripts/step_bench.py --batch 512 --seq 64 --steps 30 2>&1 | tail -1 | cut -c1-200 | sed "s/^/aten_small=$v /" | tee -a gpurun_out/final_ab.log
  done
done
