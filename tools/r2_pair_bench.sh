#!/bin/bash
mkdir -p gpurun_out/r2
for v in 1 128 0; do
DZ_CONV2D_2SM=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2/b_pair$v.json 2> gpurun_out/r2/b_pair$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/b_pair$v.json').read().strip().splitlines()[-1])
print('2SM=$v: %.1f fps, %.3f ms/step, e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
PY
done
export DZ_BATCH=8
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_pair.csv python tools/profile_frame.py > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step_pair.csv > gpurun_out/r2/pair_summary.txt; head -12 gpurun_out/r2/pair_summary.txt; tail -1 gpurun_out/r2/pair_summary.txt
