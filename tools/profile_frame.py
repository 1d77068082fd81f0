"""One eager frame of the bench workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_frame.py
(kernel launch list of exactly one warmed-up frame; numbers taken under ncu are never bench values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from detzero_b200 import synthetic as weights
from detzero_b200.det.centerpoint import build_network

sp_mode = os.environ.get('DZ_SP_MODE', weights.DEFAULT_SP_MODE)
if os.environ.get('DZ_NO_SCHEDULE'):
    from detzero_b200.spconv import pytorch as _sp
    _sp._SparseConv.SCHEDULE_TILES = False
dev = torch.device('cuda', 0)
BATCH = int(os.environ.get('DZ_BATCH', '8'))
ds, batches = bench.build_inputs(BATCH)
model = build_network(bench.make_model_cfg('VoxelBackBone8x', 'tf32', sp_mode), 3, ds).eval()
weights.load_seeded(model, 3)
bench.tune_head_for_bench(model)
model = model.to(dev)
pts = [torch.from_numpy(b['points']).to(dev) for b in batches]


def bd(i):
    b = batches[i % len(batches)]
    return {'points': pts[i % len(batches)], 'points_per_frame': b['points_per_frame'], 'frame_id': b['frame_id'], 'batch_size': b['batch_size']}


with torch.no_grad():
    for rep in range(2):
        for i in range(len(batches)):
            model(bd(i))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.forward_device(bd(0))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('done')
