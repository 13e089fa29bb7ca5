#!/bin/bash
# usage (on the GPU box, via gpurun): bash tools/gpu_cycle.sh <tag> "<test files>" [bench args...]
# tests, one default bench line, one launch list at group 32 / one lane
tag=$1; files=$2; shift 2
mkdir -p gpurun_out
FILES="$files" bash tools/gpu_run_tests.sh
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --pages 32 --e2e-pages 32 --group 32 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-iso > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log | cut -c1-200
