import ctypes, numpy as np, torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['DZ_SPCONV_DBG'] = sys.argv[1] if len(sys.argv) > 1 else '8'
from detzero_b200 import ops, _lib
from detzero_b200.spconv.pytorch import SparseConvTensor
from oracle import weights
l = ctypes.CDLL(_lib.LIB_PATH)
l.dz_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(4096, np.int64)
l.dz_debug_trace(buf.ctypes.data, 4096)     # allocate
dev = torch.device('cuda')
shape, B = [5, 188, 188], 1
idx = weights.random_sparse_coords(1, B, shape, 0.2)       # ~35K sites dense-ish
f = torch.randn(len(idx), 64)
t = SparseConvTensor(f.to(dev), torch.from_numpy(idx).to(dev), shape, B)
nbr = ops.rulebook_subm(t._idx, t._count, t._cap, t.grid_index(), [3, 3, 3])
w = torch.randn(64, 3, 3, 3, 64) * 0.05
wp = ops.pack_spconv_weight(w, _lib.DZ_TF32).to(dev)
for _ in range(3):
    out = ops.spconv_fwd(t._feat, nbr, t._count, len(idx), wp, None, None, None, True, _lib.DZ_TF32, kshape=(27, 64, 64))
torch.cuda.synchronize()
l.dz_debug_trace(buf.ctypes.data, 4096)
t0, t1, nb, t3, t4 = buf[0], buf[1], buf[2], buf[3], buf[4]
print('n sites', len(idx), 'nb', nb, 'prologue', t1 - t0, 'mainloop', t3 - t1, 'epilogue', t4 - t3, 'total', t4 - t0, 'cycles')
st = buf[8:8 + 4 * int(min(nb, 120))].reshape(-1, 4) - t0
for i in range(min(int(nb), 24)):
    print(i, 'empty_ok', st[i, 0], 'arrived', st[i, 1], 'full_ok', st[i, 2], 'mma_issued', st[i, 3])
d = np.diff(st[:, 2])
print('mean full_ok delta', d.mean(), 'median', np.median(d))
