#!/bin/bash
mkdir -p gpurun_out/r2
for h in 1 2 0; do
echo "=== DZ_CONV2D_HALO=$h"
( DZ_CONV2D_HALO=$h timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "conv2d_tf32_cta_pair" 2>&1 | tail -4 )
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 128 128 188 2>&1 | tail -2
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 256 256 94 2>&1 | tail -2
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 256 128 188 2>&1 | tail -1
done
