#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -40 ) > gpurun_out/r2/train.log; tail -40 gpurun_out/r2/train.log
