#!/bin/bash
# final validation: whole GPU suite, smoke, all bench configs on 1 GPU
mkdir -p gpurun_out/r2
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2/final_suite.log; cat gpurun_out/r2/final_suite.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/r2/final_smoke.log; cat gpurun_out/r2/final_smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/f_c2.json 2> gpurun_out/r2/f_c2.err ); echo "c2 rc=$?"
( timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2/f_c3.json 2> gpurun_out/r2/f_c3.err ); echo "c3 rc=$?"
( timeout 900 python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2/f_c4.json 2> gpurun_out/r2/f_c4.err ); echo "c4 rc=$?"
( timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/r2/f_c5.json 2> gpurun_out/r2/f_c5.err ); echo "c5 rc=$?"
( timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/f_ref.json 2> gpurun_out/r2/f_ref.err ); echo "ref rc=$?"
python - <<'PY'
import json
for c in ('c2','c3','c4','c5','ref'):
    try:
        d=json.loads([l for l in open('gpurun_out/r2/f_%s.json'%c).read().strip().splitlines() if l.startswith('{')][-1])
        print(c,'value %.1f %s ms/step %.3f'%(d['value'],d['unit'],d['ms_per_step']),'e2e',d.get('e2e') and round(d['e2e']['value'],1),'frac',d.get('roofline') and round(d['roofline']['frac'],4))
        if c=='c2':
            for k,v in d['config']['also'].items(): print('   also',k,'%.1f fps spconv %.3f ms frac %.3f'%(v['value'],v['sparse_conv_ms_per_step'],v['sparse_conv_roofline_frac']))
            print('   cpu_baseline',d.get('cpu_baseline'))
    except Exception as e: print(c,'failed',e)
PY
