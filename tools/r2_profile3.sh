#!/bin/bash
export DZ_BATCH=8
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_batch8_bf16x2_halo.csv python tools/profile_frame.py > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step_batch8_bf16x2_halo.csv --list 2>/dev/null | grep conv2d
