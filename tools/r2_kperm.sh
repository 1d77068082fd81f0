#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "spconv_fwd_fp32 or tile_schedule or frame_major or backbone3d_vs_oracle" 2>&1 | tail -3 )
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also --layer-times > gpurun_out/r2/b_kperm.json 2> gpurun_out/r2/b_kperm.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/b_kperm.json').read().strip().splitlines()[-1])
print('kperm: %.1f fps, %.3f ms/step, spconv %.3f ms frac %.4f'%(d['value'],d['ms_per_step'],d['roofline']['ms_per_step'],d['roofline']['frac']))
PY
grep "spconv layer" gpurun_out/r2/b_kperm.err | awk '{print $3,$4,$5,$6,$8,$(NF-3)}'
