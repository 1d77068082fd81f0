"""CPU restatement of the spconv 2.x semantics the reference relies on (TEST INFRASTRUCTURE).

spconv / cumm are third-party wheels absent from /root/reference and from this image, so this
file restates their published behaviour as used at
detection/detzero_det/models/centerpoint_modules/backbone3d.py:68-73,93-100,135-174,190-195 and
height_compression.py:21.  **Parity unpinned** by the reference; pinned instead against torch's dense
``F.conv3d`` on the densified input (tests/test_oracle.py) which fixes kernel-offset order, stride/padding
arithmetic, weight layout and the output-site set.

The bottom of the file is a minimal ``spconv.pytorch``-shaped shim so that the reference's *unmodified*
backbone3d.py can be imported and run on CPU by tests/golden/make_golden.py.
"""
import numpy as np
import torch
import torch.nn as nn


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


def conv_out_shape(in_shape, ksize, stride, padding, dilation=(1, 1, 1)):
    """Appendix A.2: out = floor((in + 2p - d(K-1) - 1)/s) + 1 per dim."""
    return [(in_shape[d] + 2 * padding[d] - dilation[d] * (ksize[d] - 1) - 1) // stride[d] + 1 for d in range(3)]


def _linear(idx, shape):
    """(b,z,y,x) int64 -> linear key with x fastest."""
    idx = idx.astype(np.int64)
    return ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]


def _lookup(keys_sorted, order, query):
    """index of each query key in the original (unsorted) site list, or -1."""
    pos = np.searchsorted(keys_sorted, query)
    pos = np.clip(pos, 0, len(keys_sorted) - 1)
    hit = keys_sorted[pos] == query if len(keys_sorted) else np.zeros(len(query), bool)
    return np.where(hit, order[pos], -1)


def rulebook_subm(indices, spatial_shape, ksize):
    """Submanifold rulebook (A.2): out sites == in sites (same order); pair (k, i_in, i_out) exists when
    coord[i_in] == coord[i_out] + (k - (K-1)/2).  Returns list over k=(kz*KH+ky)*KW+kx of (in_idx, out_idx)."""
    ks = _triple(ksize)
    idx = np.asarray(indices, dtype=np.int64)
    keys = _linear(idx, spatial_shape)
    order = np.argsort(keys, kind='stable')
    keys_sorted = keys[order]
    pairs = []
    n = idx.shape[0]
    out_ids = np.arange(n, dtype=np.int64)
    for kz in range(ks[0]):
        for ky in range(ks[1]):
            for kx in range(ks[2]):
                off = np.array([0, kz - (ks[0] - 1) // 2, ky - (ks[1] - 1) // 2, kx - (ks[2] - 1) // 2])
                q = idx + off
                ok = np.ones(n, bool)
                for d in range(3):
                    ok &= (q[:, d + 1] >= 0) & (q[:, d + 1] < spatial_shape[d])
                j = np.full(n, -1, np.int64)
                if ok.any():
                    j[ok] = _lookup(keys_sorted, order, _linear(q[ok], spatial_shape))
                hit = j >= 0
                pairs.append((j[hit], out_ids[hit]))
    return pairs


def rulebook_conv(indices, spatial_shape, ksize, stride, padding):
    """Regular sparse conv rulebook (A.2).  Out sites = {(i + p - k)/s : divisible, in range}; returned sorted
    ascending by linearised (b,z,y,x) (spconv's sort-unique path; order otherwise implementation-defined).
    Returns out_indices (M,4) int32, out_shape, pairs list over k of (in_idx, out_idx)."""
    ks, st, pd = _triple(ksize), _triple(stride), _triple(padding)
    idx = np.asarray(indices, dtype=np.int64)
    out_shape = conv_out_shape(list(spatial_shape), ks, st, pd)
    n = idx.shape[0]
    cand_key, cand_in, cand_k = [], [], []
    in_ids = np.arange(n, dtype=np.int64)
    k = 0
    for kz in range(ks[0]):
        for ky in range(ks[1]):
            for kx in range(ks[2]):
                kk = (kz, ky, kx)
                ok = np.ones(n, bool)
                o = np.zeros((n, 4), np.int64)
                o[:, 0] = idx[:, 0]
                for d in range(3):
                    num = idx[:, d + 1] + pd[d] - kk[d]
                    ok &= (num >= 0) & (num % st[d] == 0)
                    od = num // st[d]
                    ok &= od < out_shape[d]
                    o[:, d + 1] = od
                cand_key.append(_linear(o[ok], out_shape))
                cand_in.append(in_ids[ok])
                cand_k.append(np.full(int(ok.sum()), k, np.int64))
                k += 1
    key = np.concatenate(cand_key) if cand_key else np.zeros(0, np.int64)
    cin = np.concatenate(cand_in)
    ck = np.concatenate(cand_k)
    uniq, inv = np.unique(key, return_inverse=True)           # sorted ascending
    m = uniq.shape[0]
    out_idx = np.zeros((m, 4), np.int64)
    r = uniq.copy()
    out_idx[:, 3] = r % out_shape[2]; r //= out_shape[2]
    out_idx[:, 2] = r % out_shape[1]; r //= out_shape[1]
    out_idx[:, 1] = r % out_shape[0]; r //= out_shape[0]
    out_idx[:, 0] = r
    pairs = []
    for kk in range(ks[0] * ks[1] * ks[2]):
        sel = ck == kk
        pairs.append((cin[sel], inv[sel].astype(np.int64)))
    return out_idx.astype(np.int32), out_shape, pairs


def sparse_conv_native(features, weight, pairs, n_out, bias=None):
    """spconv CPU 'Native' algorithm (A.4): per offset gather rows -> GEMM -> scatter-add.
    features (N,Cin) f32 torch; weight (Cout,KD,KH,KW,Cin) (A.3 layout)."""
    cout = weight.shape[0]
    cin = weight.shape[-1]
    w = weight.reshape(cout, -1, cin)                        # (Cout, K, Cin)
    out = torch.zeros((n_out, cout), dtype=features.dtype)
    for k, (i_in, i_out) in enumerate(pairs):
        if len(i_in) == 0:
            continue
        g = features[torch.from_numpy(np.asarray(i_in))]
        out.index_add_(0, torch.from_numpy(np.asarray(i_out)), g @ w[:, k, :].t())
    if bias is not None:
        out = out + bias
    return out


def dense_from_sparse(features, indices, spatial_shape, batch_size):
    """SparseConvTensor.dense(): (B, C, D, H, W), zeros elsewhere (A.2)."""
    c = features.shape[1]
    out = torch.zeros((batch_size, c, *spatial_shape), dtype=features.dtype)
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    out[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = features
    return out


# ----------------------------------------------------------------------------------------------------------
# spconv.pytorch-shaped shim (CPU).  Only what backbone3d.py touches.
# ----------------------------------------------------------------------------------------------------------
class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, new_features):
        t = SparseConvTensor(new_features, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)
        return t

    def dense(self):
        return dense_from_sparse(self.features, self.indices.numpy(), self.spatial_shape, self.batch_size)


class SparseModule(nn.Module):
    pass


class _SparseConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, subm=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x):
        key = self.indice_key
        cached = x.indice_dict.get(key) if key is not None else None
        if self.subm:
            if cached is None:
                cached = ('subm', rulebook_subm(x.indices.numpy(), x.spatial_shape, self.kernel_size))
                if key is not None:
                    x.indice_dict[key] = cached
            pairs = cached[1]
            out_f = sparse_conv_native(x.features, self.weight.detach(), pairs, x.features.shape[0],
                                       None if self.bias is None else self.bias.detach())
            return SparseConvTensor(out_f, x.indices, x.spatial_shape, x.batch_size, x.indice_dict)
        if cached is None:
            out_idx, out_shape, pairs = rulebook_conv(x.indices.numpy(), x.spatial_shape, self.kernel_size,
                                                      self.stride, self.padding)
            cached = ('conv', out_idx, out_shape, pairs)
            if key is not None:
                x.indice_dict[key] = cached
        _, out_idx, out_shape, pairs = cached
        out_f = sparse_conv_native(x.features, self.weight.detach(), pairs, out_idx.shape[0],
                                   None if self.bias is None else self.bias.detach())
        return SparseConvTensor(out_f, torch.from_numpy(out_idx), out_shape, x.batch_size, x.indice_dict)


class SubMConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, 1, padding, dilation, groups, bias, indice_key, True)


class SparseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key, False)


class SparseInverseConv3d(SparseModule):
    def __init__(self, *a, **kw):
        raise NotImplementedError('never constructed by the shipped configs (backbone3d.py:72-73)')


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            else:                                             # plain nn.Module acts on .features (A.2)
                x = x.replace_feature(m(x.features))
        return x


def install_shim():
    """Put this shim at sys.modules['spconv'] / ['spconv.pytorch'] so reference files import it."""
    import sys
    import types
    pkg = types.ModuleType('spconv')
    pt = types.ModuleType('spconv.pytorch')
    for name in ('SparseConvTensor', 'SparseModule', 'SubMConv3d', 'SparseConv3d', 'SparseInverseConv3d',
                 'SparseSequential'):
        setattr(pt, name, globals()[name])
    pkg.pytorch = pt
    sys.modules['spconv'] = pkg
    sys.modules['spconv.pytorch'] = pt
    return pt
