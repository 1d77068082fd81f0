#!/bin/bash
# round-2 GPU call 1: parity suite twice (intermittent check), smoke, bench at reference precision + tf32
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2/suite1.log
( timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bench_config.py 2>&1 | tail -8 ) > gpurun_out/r2/suite2.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/r2/smoke.log
( timeout 600 python bench.py --sp-mode tf32x3 --steps 10 --warmup 3 --no-cpu-baseline --layer-times > gpurun_out/r2/bench_tf32x3.json 2> gpurun_out/r2/bench_tf32x3.err )
( timeout 600 python bench.py --sp-mode tf32 --steps 10 --warmup 3 --no-cpu-baseline --layer-times > gpurun_out/r2/bench_tf32.json 2> gpurun_out/r2/bench_tf32.err )
tail -3 gpurun_out/r2/suite1.log gpurun_out/r2/suite2.log gpurun_out/r2/smoke.log
