#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 200 python -m pytest tests/test_gpu_refine.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/attn3_t.log; cat gpurun_out/r2/attn3_t.log
timeout 100 python tools/trace_attention.py 2>&1 | grep kernel
B=4 timeout 100 python tools/trace_attention.py 2>&1 | grep kernel
