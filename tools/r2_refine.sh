#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_refine.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2/refine_t.log; cat gpurun_out/r2/refine_t.log
( timeout 900 python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2/b_c4b.json 2> gpurun_out/r2/b_c4b.err ); python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/b_c4b.json').read().strip().splitlines()[-1])
print('c4 1024pts: %.1f tracks/s %.3f ms/step frac %.3f | 256pts: %.1f tracks/s %.3f ms'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['config']['also']['256 pts/crop']['value'],d['config']['also']['256 pts/crop']['ms_per_step']))
PY
export DZ_QPTS=256
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_refine_q256b.csv python tools/profile_refine.py > gpurun_out/ncu_r.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_refine_q256b.csv | head -14
