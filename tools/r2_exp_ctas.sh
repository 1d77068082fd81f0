#!/bin/bash
mkdir -p gpurun_out/r2
for c in 148 140 132 120; do
DZ_SPCONV_CTAS=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2/b_ctas$c.json 2> gpurun_out/r2/b_ctas$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/b_ctas$c.json').read().strip().splitlines()[-1])
print('ctas $c: %.1f fps, %.3f ms/step, e2e %.1f, spconv %.3f ms'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['ms_per_step']))
PY
done
