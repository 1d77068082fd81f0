#!/bin/bash
mkdir -p gpurun_out/r2
export DZ_CONV2D_PERSIST=1
( timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "conv2d or deconv" 2>&1 | tail -3 )
BATCH=8 timeout 120 python tools/bench_conv2d.py 128 128 188 2>&1 | tail -2
BATCH=8 timeout 120 python tools/bench_conv2d.py 256 128 188 2>&1 | tail -1
BATCH=8 timeout 120 python tools/bench_conv2d.py 512 64 188 2>&1 | tail -1
BATCH=8 timeout 120 python tools/bench_conv2d.py 64 384 188 2>&1 | tail -1
BATCH=8 timeout 120 python tools/bench_conv2d.py 64 64 188 2>&1 | tail -1
( timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/p_c2.json 2> gpurun_out/r2/p_c2.err ); echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2/p_c2.json').read().strip().splitlines() if l.startswith('{')][-1])
print('c2 PERSIST value %.1f %s ms/step %.3f'%(d['value'],d['unit'],d['ms_per_step']),'e2e',round(d['e2e']['value'],1))
PY
