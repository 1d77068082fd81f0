#!/bin/bash
mkdir -p gpurun_out/r2
N=${1:-8}
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2/b_c2_n$N.json 2> gpurun_out/r2/b_c2_n$N.err ); echo "c2 n$N rc=$?"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config 5 --steps 3 --warmup 1 > gpurun_out/r2/b_c5_n$N.json 2> gpurun_out/r2/b_c5_n$N.err ); echo "c5 n$N rc=$?"
python - <<PY
import json
for c in (2,5):
    try:
        d=json.loads([l for l in open('gpurun_out/r2/b_c%d_n$N.json'%c).read().strip().splitlines() if l.startswith('{')][-1])
        print('config',c,'N=$N value %.1f %s ms/step %.3f'%(d['value'],d['unit'],d['ms_per_step']), d['config'].get('last_sequence'))
    except Exception as e:
        print('config',c,'failed',e)
PY
tail -5 gpurun_out/r2/b_c2_n$N.err gpurun_out/r2/b_c5_n$N.err
