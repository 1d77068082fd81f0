#!/bin/bash
# Diagnosis recipe for the KNOWN ISSUE in DESIGN.md section 7 (tests/test_gpu_det.py::test_dynamic_vfe_into_backbone fails on some
# boxes, then persistently inside that process).  Run under gpurun; everything lands in gpurun_out/.
#   1. does this box reproduce?  (prints the failing checks)
#   2. initcheck: reads of device memory no kernel / memset / copy ever wrote   3. memcheck: out-of-bounds accesses
export PYTORCH_NO_CUDA_MEMORY_CACHING=1      # one cudaMalloc per tensor: the sanitizer sees exact allocation bounds
T="tests/test_gpu_det.py -q -m gpu -k dynamic_vfe_into --runxfail -x -s"
for i in 1 2 3; do python -m pytest $T 2>&1 | grep -E "FIRST ATTEMPT|passed|failed" ; done | tee gpurun_out/hunt_repro.txt
timeout 900 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest $T > gpurun_out/hunt_initcheck.txt 2>&1
grep -E "Uninitialized|ERROR SUMMARY|at .*\.cu" gpurun_out/hunt_initcheck.txt | head -40
timeout 900 compute-sanitizer --tool memcheck --print-limit 40 python -m pytest $T > gpurun_out/hunt_memcheck.txt 2>&1
grep -E "Invalid|ERROR SUMMARY|at .*\.cu" gpurun_out/hunt_memcheck.txt | head -40
