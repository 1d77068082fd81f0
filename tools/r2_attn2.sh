#!/bin/bash
# single-pass attention: parity tests + timing A/B against the two-pass kernel
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_refine.py -x -q 2>&1 | tail -8 ) > gpurun_out/r2/attn2_t.log; cat gpurun_out/r2/attn2_t.log
for tp in 0 1; do
export DZ_ATTN_TWO_PASS=$tp
timeout 300 python - <<'PY'
import os, torch
from detzero_b200 import ops, _lib
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
B, H, dh, Pq, Pk = 4, 8, 32, 1024, 9600
q = (torch.randn(B, Pq, H*dh, generator=g)*0.5).to(dev); k = torch.randn(B, Pk, H*dh, generator=g).to(dev); v = torch.randn(B, Pk, H*dh, generator=g).to(dev)
mask = torch.zeros(B, Pk, dtype=torch.uint8); mask[:, 9000:] = 1; mask = mask.to(dev)
for _ in range(3): out = ops.attention(q, k, v, mask, H, mode=_lib.DZ_TF32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = ops.attention(q, k, v, mask, H, mode=_lib.DZ_TF32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
fl = 4.0*B*H*Pq*Pk*dh
print('two_pass=%s  %.1f us  %.1f TFLOP/s (QK+PV useful)' % (os.environ.get('DZ_ATTN_TWO_PASS'), ms*1e3, fl/ms/1e9))
PY
done
unset DZ_ATTN_TWO_PASS
( timeout 900 python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2/b_c4c.json 2> gpurun_out/r2/b_c4c.err ); python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/b_c4c.json').read().strip().splitlines()[-1])
print('c4 1024pts: %.1f tracks/s %.3f ms/step frac %.3f | 256pts: %.1f tracks/s %.3f ms'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['config']['also']['256 pts/crop']['value'],d['config']['also']['256 pts/crop']['ms_per_step']))
PY
