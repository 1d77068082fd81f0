#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2/suite_full.log
tail -5 gpurun_out/r2/suite_full.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/r2/smoke2.log; cat gpurun_out/r2/smoke2.log
