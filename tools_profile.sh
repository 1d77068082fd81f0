#!/bin/bash
# usage (under gpurun): bash tools_profile.sh <tag> [DZ_BATCH]   -> gpurun_out/launches_<tag>.csv + summary on stdout
# kernel launch list (gpu__time_duration) of exactly ONE warmed-up eager step of the bench workload
tag=$1; export DZ_BATCH=${2:-8}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_frame.py > gpurun_out/ncu_$tag.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$tag.csv | head -40
# full counters of the sparse-conv launches of the same step:
#   ncu --profile-from-start off --set full --clock-control none -k regex:k_spconv_tf32 -o gpurun_out/prof_spconv_$tag python tools/profile_frame.py
