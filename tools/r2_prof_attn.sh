#!/bin/bash
# full counters of the single-pass attention kernel inside one refiner step (16 tracks, 256 pts/crop) + the step's launch list
export DZ_QPTS=256
mkdir -p gpurun_out
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_attention_tf32_v2 -c 3 -o gpurun_out/prof_attention_v2_r2 -f python tools/profile_refine.py > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_refine_step_16tracks_256pts_v2.csv python tools/profile_refine.py > gpurun_out/ncu_r.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_refine_step_16tracks_256pts_v2.csv 2>/dev/null | head -8
