#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2/b_e2e.json 2> gpurun_out/r2/b_e2e.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2/b_e2e.json").read().strip().splitlines()[-1])
print("value %.1f ms %.3f e2e %.1f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d["e2e"])
PY
tail -3 gpurun_out/r2/b_e2e.err
