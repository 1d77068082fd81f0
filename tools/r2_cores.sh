#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "spconv_fwd_fp32 or tile_schedule" 2>&1 | tail -2 )
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2/b_cores.json 2> gpurun_out/r2/b_cores.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/b_cores.json').read().strip().splitlines()[-1])
print('coresident build: %.1f fps, %.3f ms/step, e2e %.1f, spconv %.3f ms frac %.4f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['ms_per_step'],d['roofline']['frac']))
PY
