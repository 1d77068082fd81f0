#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "deconv or conv2d" 2>&1 | tail -4 )
( timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/f_c2.json 2> gpurun_out/r2/f_c2.err ); echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/f_c2.json').read().strip().splitlines() if l.startswith('{')][-1])
print('c2 value %.1f %s ms/step %.3f'%(d['value'],d['unit'],d['ms_per_step']),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4))
PY
