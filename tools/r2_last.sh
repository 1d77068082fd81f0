#!/bin/bash
mkdir -p gpurun_out/r2
timeout 60 python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/r2/l_c4.json 2> gpurun_out/r2/l_c4.err; echo "rc=$?"; tail -c 600 gpurun_out/r2/l_c4.json
