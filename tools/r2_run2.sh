#!/bin/bash
# round-2 GPU call 2: the persistent bf16-plane sparse conv: kernel parity first, then backbone / full-lattice parity, then bench
mkdir -p gpurun_out/r2
T="tests/test_gpu_det.py"
( timeout 300 python -m pytest $T -x -q -k "planes_roundtrip or spconv_fwd_fp32 or tile_schedule or frame_major" 2>&1 | tail -25 ) > gpurun_out/r2/k1.log
tail -3 gpurun_out/r2/k1.log
if grep -q "failed\|error" gpurun_out/r2/k1.log; then exit 1; fi
( timeout 600 python -m pytest $T -x -q -k "backbone3d_vs_oracle" 2>&1 | tail -25 ) > gpurun_out/r2/k2.log
tail -3 gpurun_out/r2/k2.log
( timeout 900 python -m pytest tests/test_gpu_bench_config.py -q 2>&1 | tail -40 ) > gpurun_out/r2/k3.log
tail -5 gpurun_out/r2/k3.log
for m in bf16x2 bf16 tf32; do
( timeout 600 python bench.py --sp-mode $m --steps 10 --warmup 3 --no-cpu-baseline --layer-times > gpurun_out/r2/bench2_$m.json 2> gpurun_out/r2/bench2_$m.err )
tail -c 600 gpurun_out/r2/bench2_$m.json
done
