#!/bin/bash
# usage (GPU box): bash tools/gpu_profile.sh <tag> — one group of 32 sheets on one lane: the plain run first, then the
# launch list, then ONE ncu --set full capture of the group's launches (+ its raw page as csv)
tag=$1
mkdir -p gpurun_out
ARGS="--pages 32 --e2e-pages 32 --group 32 --lanes 1 --steps 1 --warmup 3 --no-cpu-baseline --no-iso"
python bench.py $ARGS > gpurun_out/pre_$tag.json 2> gpurun_out/pre_$tag.err || { tail -5 gpurun_out/pre_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -c 46 -f -o gpurun_out/prof_$tag python bench.py $ARGS > gpurun_out/prof_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2> /dev/null
ls -la gpurun_out/prof_$tag*
tail -2 gpurun_out/prof_$tag.log | cut -c1-200
