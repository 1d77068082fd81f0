#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 240 python -m pytest tests/test_gpu_det.py -x -q -k "cta_pair" 2>&1 | tail -15 ) > gpurun_out/r2/pair.log; cat gpurun_out/r2/pair.log
