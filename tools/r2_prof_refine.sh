#!/bin/bash
for q in 256 1024; do
export DZ_QPTS=$q
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_refine_q$q.csv python tools/profile_refine.py > gpurun_out/ncu_r.log 2>&1
echo "=== qpts $q"; python tools/summarize_launches.py gpurun_out/r02_launches_refine_q$q.csv --list | head -80
done
