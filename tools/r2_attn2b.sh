#!/bin/bash
mkdir -p gpurun_out/r2
timeout 300 python tools/trace_attention.py 2>&1 | tail -8
B=4 timeout 300 python tools/trace_attention.py 2>&1 | tail -6
