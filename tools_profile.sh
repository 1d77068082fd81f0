#!/bin/bash
# usage (under gpurun): bash tools_profile.sh <tag> [bench args...]   -> gpurun_out/launches_<tag>.csv
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-450} -c ${NCU_COUNT:-120} --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log | cut -c1-200
