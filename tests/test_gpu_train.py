"""Training path: sparse-conv dgrad / wgrad / bias grad, train-mode BatchNorm and SparseInverseConv3d against
PyTorch autograd on the densified problem (the dependency-free equivalent SURVEY §8c recommends): densify, F.conv3d /
F.conv_transpose3d, read the result at the sparse sites, back-propagate a random cotangent."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_ref, weights
from tests import util

pytestmark = pytest.mark.gpu


def _dense(idx, f, shape, B):
    d = torch.zeros((B, f.shape[1], *shape), dtype=torch.float64)
    d[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = f.double()
    return d


@pytest.mark.parametrize('subm,ks,stride,pad,cin,cout', [(True, 3, 1, 1, 16, 32), (False, 3, 2, 1, 32, 64), (False, (3, 1, 1), (2, 1, 1), 0, 64, 128),
                                                         (True, 3, 1, 1, 128, 128)])
def test_sparse_conv_backward_vs_dense_autograd(cuda, subm, ks, stride, pad, cin, cout):
    from detzero_b200.spconv import pytorch as sp
    from detzero_b200.spconv.pytorch import _triple
    shape, B = [9, 20, 18], 2
    idx = weights.random_sparse_coords(5, B, shape, 0.12)
    g = torch.Generator().manual_seed(3)
    f = torch.randn(len(idx), cin, generator=g)
    conv = (sp.SubMConv3d if subm else sp.SparseConv3d)(cin, cout, ks, stride=stride, padding=pad, bias=True, indice_key='k').to(cuda).train()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x_feat = f.to(cuda).requires_grad_(True)
    x = sp.SparseConvTensor(x_feat, torch.from_numpy(idx).to(cuda), shape, B)
    y = conv(x)
    out_idx = y.indices.cpu().numpy()
    cot = torch.randn(y.features.shape, generator=g)
    (y._feat * cot.to(cuda)).sum().backward()
    # dense reference in float64
    ksz, st, pd = _triple(ks), _triple(stride), _triple(pad if not subm else 1)
    xd = _dense(idx, f, shape, B).requires_grad_(True)
    wd = conv.weight.detach().cpu().double().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)       # (cout, cin, kd, kh, kw)
    bd = conv.bias.detach().cpu().double().requires_grad_(True)
    yd = F.conv3d(xd, wd, bd, stride=st, padding=pd)
    if subm:
        assert np.array_equal(out_idx, idx)
    else:
        ref_idx, _, _ = spconv_ref.rulebook_conv(idx, shape, ks, stride, pad)
        assert np.array_equal(out_idx, ref_idx)
    ys = yd[out_idx[:, 0], :, out_idx[:, 1], out_idx[:, 2], out_idx[:, 3]]
    assert util.rel_err(y.features.detach().cpu(), ys.detach()) < 1e-5
    (ys * cot.double()).sum().backward()
    gin = xd.grad[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    assert util.rel_err(x_feat.grad.cpu(), gin) < 1e-5
    assert util.rel_err(conv.weight.grad.cpu(), wd.grad.permute(0, 2, 3, 4, 1)) < 1e-5
    assert util.rel_err(conv.bias.grad.cpu(), bd.grad) < 1e-5


def test_inverse_conv_vs_dense_transposed_conv(cuda):
    """SparseInverseConv3d: output sites = the matching SparseConv3d's input sites; values = the dense transposed conv read there"""
    from detzero_b200.spconv import pytorch as sp
    shape, B, cin, cmid, cout = [9, 20, 18], 2, 16, 32, 16
    idx = weights.random_sparse_coords(8, B, shape, 0.1)
    g = torch.Generator().manual_seed(4)
    f = torch.randn(len(idx), cin, generator=g)
    down = sp.SparseConv3d(cin, cmid, 3, stride=2, padding=1, bias=False, indice_key='sp').to(cuda).train()
    up = sp.SparseInverseConv3d(cmid, cout, 3, indice_key='sp', bias=False).to(cuda).train()
    with torch.no_grad():
        down.weight.copy_(torch.randn(down.weight.shape, generator=g) * 0.1)
        up.weight.copy_(torch.randn(up.weight.shape, generator=g) * 0.1)
    x = sp.SparseConvTensor(f.to(cuda), torch.from_numpy(idx).to(cuda), shape, B)
    mid = down(x)
    mid_feat = mid._feat.detach().clone().requires_grad_(True)
    y = up(mid.replace_feature(mid_feat))
    assert np.array_equal(y.indices.cpu().numpy(), idx) and y.spatial_shape == shape
    cot = torch.randn(y.features.shape, generator=g)
    (y._feat * cot.to(cuda)).sum().backward()
    mi = mid.indices.cpu().numpy()
    md = _dense(mi, mid_feat.detach().cpu(), mid.spatial_shape, B).requires_grad_(True)
    wt = up.weight.detach().cpu().double().permute(4, 0, 1, 2, 3).contiguous().requires_grad_(True)           # (cin, cout, kd, kh, kw)
    full = F.conv_transpose3d(md, wt, stride=2, padding=1, output_padding=1)
    yd = full[:, :, :shape[0], :shape[1], :shape[2]]
    ys = yd[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    assert util.rel_err(y.features.detach().cpu(), ys.detach()) < 1e-5
    (ys * cot.double()).sum().backward()
    assert util.rel_err(mid_feat.grad.cpu(), md.grad[mi[:, 0], :, mi[:, 1], mi[:, 2], mi[:, 3]]) < 1e-5
    assert util.rel_err(up.weight.grad.cpu(), wt.grad.permute(1, 2, 3, 4, 0)) < 1e-5


@pytest.mark.parametrize('kind', ['VoxelBackBone8x', 'VoxelResBackBone8x'])
def test_backbone_trains(cuda, kind):
    """train mode: BatchNorm1d uses batch statistics over the valid rows, every parameter receives a finite gradient, an SGD step
    lowers the loss; eval mode afterwards still runs the fused path"""
    from detzero_b200.det import cp_modules
    cfg = util.model_cfg(kind).BACKBONE_3D
    m = cp_modules[kind](model_cfg=cfg, input_channels=5, grid_size=[96, 96, 40])
    weights.load_seeded(m, 5)
    m = m.to(cuda).train()
    g = np.random.default_rng(0)
    pts = util.clustered_cloud(6000, 3)
    pts = pts[(np.abs(pts[:, 0]) < 4.8) & (np.abs(pts[:, 1]) < 4.8)]
    coords = np.unique(np.floor((pts[:, :3] - np.array([-4.8, -4.8, -2])) / np.array(util.VOXEL)).astype(np.int32)[:, ::-1], axis=0)
    coords = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], 1)
    feats = torch.from_numpy(g.normal(0, 1, (len(coords), 5)).astype(np.float32)).to(cuda)
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    target = None
    losses = []
    for it in range(3):
        bd = m({'voxel_features': feats, 'voxel_coords': torch.from_numpy(coords).to(cuda), 'batch_size': 1})
        out = bd['encoded_spconv_tensor']
        if target is None:
            target = torch.randn_like(out._feat)
        loss = ((out._feat - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        for n_, p in m.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), n_
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    # train-mode BN == torch BN on the valid rows (single layer check)
    seq = m.conv_input
    x = seq(m._input_tensor({'voxel_features': feats, 'voxel_coords': torch.from_numpy(coords).to(cuda), 'batch_size': 1}))
    assert x._feat.shape[0] == len(coords) and torch.isfinite(x._feat).all()
    m.eval()
    with torch.no_grad():
        bd = m({'voxel_features': feats, 'voxel_coords': torch.from_numpy(coords).to(cuda), 'batch_size': 1})
    assert torch.isfinite(bd['encoded_spconv_tensor'].features).all()
