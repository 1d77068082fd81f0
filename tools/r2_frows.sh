#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_track.py tests/test_gpu_io.py -x -q 2>&1 | tail -30 ) > gpurun_out/r2/frows.log; tail -30 gpurun_out/r2/frows.log
