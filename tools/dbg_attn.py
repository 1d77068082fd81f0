import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detzero_b200 import ops, _lib
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
for (B, Pq, Pk, masked) in [(1, 128, 128, False), (1, 128, 256, False), (3, 200, 200, True), (3, 200, 9600, True)]:
    H, dh = 8, 32
    q = torch.randn(B, Pq, H * dh, generator=g) * 0.5
    k = torch.randn(B, Pk, H * dh, generator=g)
    v = torch.randn(B, Pk, H * dh, generator=g)
    mask = torch.zeros(B, Pk, dtype=torch.bool)
    if masked:
        for b in range(B):
            mask[b, int(Pk * (0.3 + 0.3 * b)):] = True
    qh, kh, vh = [t.view(B, -1, H, dh).permute(0, 2, 1, 3) for t in (q, k, v)]
    s = (qh @ kh.transpose(-1, -2)).masked_fill(mask[:, None, None, :], float('-inf'))
    ref = (torch.softmax(s, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(B, Pq, H * dh)
    out = ops.attention(q.to(dev), k.to(dev), v.to(dev), mask.to(torch.uint8).to(dev) if masked else None, H, mode=_lib.DZ_TF32).cpu()
    err = (out - ref).abs()
    print((B, Pq, Pk, masked), 'rel err %.3e' % (err.max() / ref.abs().max()).item(), 'nan', torch.isnan(out).sum().item(),
          'bad rows', (err.max(dim=2)[0] > 1e-2).sum().item(), 'of', B * Pq, 'bad per batch', [(err[b].max().item()) for b in range(B)])
    if err.max() > 1e-2:
        bad = (err > 1e-2).nonzero()
        print('  first bad', bad[:5].tolist(), 'last bad', bad[-3:].tolist())
        # per (head) error and per dim pattern
        e2 = err.view(B, Pq, H, dh)
        print('  per-head max', e2.amax(dim=(0, 1, 3)).tolist())
        print('  per-dim max', [round(x, 3) for x in e2.amax(dim=(0, 1, 2)).tolist()])
