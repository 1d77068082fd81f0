"""Per-role wait / work clock accounting of CTA 0 of the persistent bf16-plane sparse conv, for every sparse-conv launch of one
step of the bench workload.  usage: python tools/trace_spconv_bf16.py [planes=2] [batch=8]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from detzero_b200 import ops, _lib, synthetic
from detzero_b200.det import build_network
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = 'bf16x2' if P == 2 else 'bf16'
dev = torch.device('cuda')
ds, batches = bench.build_inputs(batch)
model = build_network(bench.make_model_cfg('VoxelBackBone8x', 'tf32', mode), 3, ds).eval()
synthetic.load_seeded(model, 3)
model = model.to(dev)


def bd(i):
    b = batches[i % len(batches)]
    return {'points': torch.from_numpy(b['points']).to(dev), 'points_per_frame': b['points_per_frame'], 'frame_id': b['frame_id'],
            'batch_size': b['batch_size']}


for i in range(4):
    with torch.no_grad():
        try:
            model(bd(i))
        except RuntimeError as e:
            print('settle:', str(e)[:80])
l = ctypes.CDLL(_lib.LIB_PATH)
l.dz_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(4096, np.int64)
l.dz_debug_trace(buf.ctypes.data, 4096)          # allocate + switch the trace on
orig = ops.spconv_fwd
rows = []


def traced(*a, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(*a, **kw)
    e1.record()
    torch.cuda.synchronize()
    l.dz_debug_trace(buf.ctypes.data, 4096)
    rows.append((kw.get('kshape'), e0.elapsed_time(e1) * 1000, buf[2048:2048 + 32].copy()))
    return out


ops.spconv_fwd = traced
from detzero_b200.spconv import pytorch as sp
with torch.no_grad():
    model.forward_device(bd(0))
torch.cuda.synchronize()
print('mode %s batch %d | clocks of CTA 0' % (mode, batch))
print('%-14s %7s | %9s | tile: %8s %5s | prod0: %8s %8s %8s %6s | mma: %8s %8s %8s %6s | epi0: %8s %8s %8s %5s' % (
    'K,cin,cout', 'us', 'total', 'w_empty', 'tiles', 'w_nbr', 'w_empty', 'issue', 'steps', 'w_nbr+acc', 'w_full', 'issue', 'steps',
    'w_nbr', 'w_accfull', 'body', 'tiles'))
for ks, us, d in rows:
    print('%-14s %7.1f | %9d | tile: %8d %5d | prod0: %8d %8d %8d %6d | mma: %8d %8d %8d %6d | epi0: %8d %8d %8d %5d' % (
        str(ks), us, d[20], d[0], d[1], d[8], d[9], d[10], d[11], d[16], d[17], d[18], d[19], d[24], d[25], d[26], d[27]))
