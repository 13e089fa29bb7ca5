#!/bin/bash
# usage: bash tools/gpu_run_tests.sh [pytest -k expression] — GPU parity suite, one process per file
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${FILES:-test_gpu_blit test_gpu_output test_gpu_filters test_gpu_fuzz test_gpu_engine}; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout=240 --timeout-method=thread -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
