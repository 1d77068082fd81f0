"""One eager refiner step (PRM + GRM, 16 tracks) between cudaProfilerStart/Stop for an ncu launch list.  DZ_QPTS=256|1024"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
env = bench.Env()
rf = bench.Refiner(env, os.environ.get('DZ_MODE', 'tf32'), int(os.environ.get('DZ_TRACKS', '16')), int(os.environ.get('DZ_QPTS', '256')),
                   with_grm=not os.environ.get('DZ_NO_GRM'))
for i in range(3):
    rf.step_resident(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rf.step_resident(0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
