#!/bin/bash
# one eager 8-frame step of the default bench workload under ncu: launch list, then full counters of the two conv kernel families
export DZ_BATCH=8
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_batch8.csv python tools/profile_frame.py > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step_batch8.csv | head -40
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_bf16 -o gpurun_out/prof_spconv_r2 -f python tools/profile_frame.py > gpurun_out/ncu_s.log 2>&1; tail -2 gpurun_out/ncu_s.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv2d_tf32 -c 22 -o gpurun_out/prof_conv2d_r2 -f python tools/profile_frame.py > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep
