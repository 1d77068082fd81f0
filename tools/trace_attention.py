"""Clock accounting of CTA (0,0) of k_attention_tf32 on the PRM cross-attention shape (16 tracks, 200 queries, 9600 keys, 8 heads)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detzero_b200 import ops, _lib
l = ctypes.CDLL(_lib.LIB_PATH)
l.dz_debug_attention_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, Pq, Pk, H, E = int(os.environ.get('B', 16)), 200, int(os.environ.get('PK', 9600)), 8, 256
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
q = (torch.randn(B, Pq, E, generator=g) * 0.5).to(dev)
kv = torch.randn(B, Pk, 2 * E, generator=g).to(dev)
k, v = kv[:, :, :E], kv[:, :, E:]
mask = torch.zeros(B, Pk, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.attention(q, k, v, mask, H, mode=_lib.DZ_TF32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(q, k, v, mask, H, mode=_lib.DZ_TF32); e1.record(); torch.cuda.synchronize()
print('kernel %.1f us' % (1000 * e0.elapsed_time(e1)))
l.dz_debug_attention_trace(None, 1)
ops.attention(q, k, v, mask, H, mode=_lib.DZ_TF32)
buf = np.zeros(64, np.int64)
l.dz_debug_attention_trace(buf.ctypes.data, 0)
n = buf[24]
print('blocks', n)
if not os.environ.get('DZ_ATTN_TWO_PASS'):
    print('loader : wait k_empty %d, wait v_empty %d, V^T work %d, total %d' % tuple(buf[0:4]))
    print('mma    : wait k_full %d, wait s_empty %d, wait v_full %d, wait p_full %d, wait o_empty %d | total %d' % (tuple(buf[8:13]) + (buf[14],)))
    print('softmax: wait s_full %d, ld S + max %d, bar.sync %d, exp %d, wait p_empty %d, write P %d, wait o_full %d, fold %d | total %d' % (tuple(buf[16:24]) + (buf[25],)))
    sys.exit(0)
print('loader : wait k_empty %d, wait v_empty %d, V^T work %d, total %d' % tuple(buf[0:4]))
print('mma    : wait k_full %d, wait s_empty %d, wait v_full %d, wait p_full %d, issue S %d, issue PV %d | pass1 end %d, total %d' % (tuple(buf[8:14]) + (buf[15], buf[14])))
print('softmax: p1 wait s_full %d, p1 body %d | p2 wait s_full %d, wait p_empty %d, body %d | pass1 end %d total %d' % (tuple(buf[16:21]) + (buf[23], buf[22])))
