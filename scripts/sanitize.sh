#!/usr/bin/env bash
# Race / memory checking of the single-GPU kernels: compute-sanitizer memcheck + racecheck over the
# op tests (tiny shapes).  Run on a GPU box:  gpurun -- 'bash scripts/sanitize.sh'
set -uo pipefail
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -x -q \
      -k "embed or norm or colsum or adamw or weighted_avg or checksum or meta_kernels" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitizer_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3 | tee -a gpurun_out/sanitizer_summary.txt
done
