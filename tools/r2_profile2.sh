#!/bin/bash
# after the halo conv kernel: launch list of one eager 8-frame step + full counters of the dense conv kernels
export DZ_BATCH=8
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_batch8_bf16x2_halo.csv python tools/profile_frame.py > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step_batch8_bf16x2_halo.csv | head -24
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv2d_tf32 -c 24 -o gpurun_out/prof_conv2d_halo_r2 -f python tools/profile_frame.py > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep
