#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_refine.py -x -q 2>&1 | tail -8 ) > gpurun_out/r2/attn2_t.log; cat gpurun_out/r2/attn2_t.log
timeout 300 python tools/trace_attention.py 2>&1 | tail -5
B=4 timeout 300 python tools/trace_attention.py 2>&1 | grep kernel
DZ_ATTN_TWO_PASS=1 timeout 300 python tools/trace_attention.py 2>&1 | grep kernel
( timeout 900 python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2/b_c4c.json 2> gpurun_out/r2/b_c4c.err ); python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/b_c4c.json').read().strip().splitlines()[-1])
print('c4 1024pts: %.1f tracks/s %.3f ms/step frac %.3f | 256pts: %.1f tracks/s %.3f ms'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['config']['also']['256 pts/crop']['value'],d['config']['also']['256 pts/crop']['ms_per_step']))
PY
