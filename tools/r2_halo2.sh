#!/bin/bash
for h in 1 2; do
echo "=== DZ_CONV2D_HALO=$h"
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 512 64 188 2>&1 | tail -1
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 64 384 188 2>&1 | tail -1
DZ_CONV2D_HALO=$h BATCH=8 timeout 120 python tools/bench_conv2d.py 64 64 188 2>&1 | tail -2
done
( DZ_CONV2D_HALO=2 timeout 300 python -m pytest tests/test_gpu_det.py -x -q -k "conv2d_tf32_cta_pair" 2>&1 | tail -2 )
