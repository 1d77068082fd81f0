#!/bin/bash
mkdir -p gpurun_out/r2
T="tests/test_gpu_det.py"
( timeout 300 python -m pytest $T -x -q -k "planes_roundtrip or spconv_fwd_fp32 or tile_schedule or frame_major" 2>&1 | tail -25 ) > gpurun_out/r2/k1.log
tail -3 gpurun_out/r2/k1.log
if grep -q "failed\|error" gpurun_out/r2/k1.log; then exit 1; fi
( timeout 600 python -m pytest $T -x -q -k "backbone3d_vs_oracle and bf16" 2>&1 | tail -25 ) > gpurun_out/r2/k2.log
tail -3 gpurun_out/r2/k2.log
for m in bf16x2 bf16; do
( timeout 600 python bench.py --sp-mode $m --steps 10 --warmup 3 --no-cpu-baseline --layer-times > gpurun_out/r2/bench3_$m.json 2> gpurun_out/r2/bench3_$m.err )
tail -c 500 gpurun_out/r2/bench3_$m.json; grep "spconv layer" gpurun_out/r2/bench3_$m.err | awk '{print $3,$4,$5,$6,$8,$(NF-1)}'
done
export DZ_NO_FRAME_MAJOR=1
( timeout 600 python bench.py --sp-mode bf16x2 --steps 10 --warmup 3 --no-cpu-baseline --layer-times > gpurun_out/r2/bench3_nofm.json 2> gpurun_out/r2/bench3_nofm.err )
tail -c 500 gpurun_out/r2/bench3_nofm.json; grep "spconv layer" gpurun_out/r2/bench3_nofm.err | awk '{print $3,$4,$5,$6,$8,$(NF-1)}'
