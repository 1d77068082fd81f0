#!/bin/bash
# 2-GPU call: NCCL gather test, default bench at N=2 (gather in the timed region), config 5 at N=2
mkdir -p gpurun_out/r2
( timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2/nccl_test.log; cat gpurun_out/r2/nccl_test.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/b_c2_n2.json 2> gpurun_out/r2/b_c2_n2.err ); echo "c2 n2 rc=$?"; tail -c 900 gpurun_out/r2/b_c2_n2.json | head -c 600; tail -3 gpurun_out/r2/b_c2_n2.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --steps 3 --warmup 1 > gpurun_out/r2/b_c5_n2.json 2> gpurun_out/r2/b_c5_n2.err ); echo "c5 n2 rc=$?"; tail -c 1500 gpurun_out/r2/b_c5_n2.json; tail -3 gpurun_out/r2/b_c5_n2.err
