#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/b_c2.json 2> gpurun_out/r2/b_c2.err ); echo "c2 rc=$?"; tail -c 1500 gpurun_out/r2/b_c2.json; tail -3 gpurun_out/r2/b_c2.err
( timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2/b_c3.json 2> gpurun_out/r2/b_c3.err ); echo "c3 rc=$?"; tail -c 800 gpurun_out/r2/b_c3.json; tail -3 gpurun_out/r2/b_c3.err
( timeout 900 python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2/b_c4.json 2> gpurun_out/r2/b_c4.err ); echo "c4 rc=$?"; tail -c 1200 gpurun_out/r2/b_c4.json; tail -3 gpurun_out/r2/b_c4.err
( timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/r2/b_c5.json 2> gpurun_out/r2/b_c5.err ); echo "c5 rc=$?"; tail -c 1200 gpurun_out/r2/b_c5.json; tail -3 gpurun_out/r2/b_c5.err
