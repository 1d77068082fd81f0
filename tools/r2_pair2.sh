#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_det.py tests/test_gpu_bench_config.py -x -q -k "conv2d or bev or centerpoint or batched or full_lattice" 2>&1 | tail -5 ) > gpurun_out/r2/pair2.log; cat gpurun_out/r2/pair2.log
for v in 1 0; do
DZ_CONV2D_2SM=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2/b_pairb$v.json 2> gpurun_out/r2/b_pairb$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/b_pairb$v.json').read().strip().splitlines()[-1])
print('2SM=$v: %.1f fps, %.3f ms/step, e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
PY
done
