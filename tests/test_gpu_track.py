"""GPU parity tests for the steps around the detector / refiner: tracker association matrices and the
object crop, against the C / numpy oracle restatements of the reference's iou3d_nms and roiaware_pool3d code."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _boxes(g, n):
    b = np.concatenate([g.uniform(-20, 20, (n, 2)), g.uniform(-1, 1, (n, 1)), g.uniform(1.5, 5, (n, 2)), g.uniform(1, 2, (n, 1)),
                        g.uniform(-3.2, 3.2, (n, 1))], axis=1).astype(np.float32)
    m = n // 3
    b[m:2 * m, :2] = b[:m, :2] + g.normal(0, 0.3, (m, 2)).astype(np.float32)          # overlapping clusters (what the tracker associates)
    b[m:2 * m, 3:7] = b[:m, 3:7] + g.normal(0, 0.05, (m, 4)).astype(np.float32)
    return b


def test_tracker_distance_matrices(cuda):
    """IoUBEV / BEV overlap / IoU3D / IoU2D between two frames' boxes, read IN PLACE from the gathered (F, 500, 9) tensor"""
    from detzero_b200 import track
    g = np.random.default_rng(4)
    gathered = np.zeros((3, 500, 9), np.float32)
    n0, n1 = 210, 187
    gathered[0, :n0, :7], gathered[1, :n1, :7] = _boxes(g, n0), _boxes(g, n1)
    gathered[1, :60, :7] = gathered[0, :60, :7] + g.normal(0, 0.1, (60, 7)).astype(np.float32)      # same objects one frame later
    t = torch.from_numpy(gathered).to(cuda)
    a, b = t[0, :n0], t[1, :n1]                              # strided views: row stride 9 floats, no copy
    A, B = gathered[0, :n0, :7], gathered[1, :n1, :7]
    iou = track.IoUBEV_dis_mat(a, b).cpu().numpy()
    assert np.abs(iou - oracle.boxes_iou_bev(A, B)).max() < 1e-4
    ov = track.bev_overlap_gpu(a, b).cpu().numpy()
    assert np.abs(ov - oracle.boxes_overlap_bev(A, B)).max() < 2e-3            # areas up to ~25 m^2
    iou3 = track.IoU3D_dis_mat(a, b).cpu().numpy()
    assert np.abs(iou3 - oracle.boxes_iou3d(A, B)).max() < 1e-4
    xywh_a, xywh_b = A[:, [0, 1, 3, 4]], B[:, [0, 1, 3, 4]]
    iou2 = track.IoU2D_dis_mat(torch.from_numpy(xywh_a).to(cuda), torch.from_numpy(xywh_b).to(cuda)).cpu().numpy()
    assert np.abs(iou2 - oracle.iou2d(xywh_a, xywh_b)).max() < 1e-6
    assert (iou > 0.5).sum() >= 40                                              # the matrices are not trivially empty
    e = track.IoUBEV_dis_mat(a[:0], b)                                          # empty side -> (0, M) like the reference
    assert e.shape == (0, n1)


def test_object_crop_vs_oracle(cuda):
    """points_in_boxes_gpu_v2 + the daemon's ordered per-object crop: mask == oracle except for points within 1e-4 m of a box face
    (CPU vs GPU cosf/sinf), ordered indices == flatnonzero of the mask, counts exact, cap respected"""
    from detzero_b200 import track
    from detzero_b200.det.dataset import synth_waymo_cloud
    g = np.random.default_rng(9)
    pts = synth_waymo_cloud(3, 60000)[:, :4]
    T = 40
    boxes = np.zeros((T, 9), np.float32)
    centres = pts[g.integers(0, pts.shape[0], T), :3]
    boxes[:, :3] = centres + g.normal(0, 0.2, (T, 3)).astype(np.float32)
    boxes[:, 3:6] = g.uniform([3, 1.5, 1.2], [8, 3, 3], (T, 3)) * 1.1            # enlarged like the daemon (enlarge_scale)
    boxes[:, 6] = g.uniform(-3.2, 3.2, T)
    boxes[5, 5] = 100.0                                                           # crop_on_bev: dz = 100
    want = oracle.points_in_boxes(pts[:, :3], boxes[:, :7])
    p_d, b_d = torch.from_numpy(pts).to(cuda), torch.from_numpy(boxes).to(cuda)
    mask = track.points_in_boxes_gpu_v2(p_d[None, :, :3].contiguous(), b_d[None, :, :7].contiguous())[0].cpu().numpy()
    diff = np.argwhere(mask != want)
    for t, m in diff:                                                             # only points ON a face may differ
        b, q = boxes[t].astype(np.float64), pts[m].astype(np.float64)
        c, s = np.cos(-b[6]), np.sin(-b[6])
        lx, ly = (q[0] - b[0]) * c - (q[1] - b[1]) * s, (q[0] - b[0]) * s + (q[1] - b[1]) * c
        d = min(abs(abs(lx) - b[3] / 2), abs(abs(ly) - b[4] / 2), abs(abs(q[2] - b[2]) - b[5] / 2))
        assert d < 1e-4, (t, m, d)
    assert len(diff) <= 5 and want.sum() > 2000
    cap = 256
    idx, num = track.crop_points_in_boxes(p_d, b_d, cap)                          # points (M,4) stride 4, boxes (T,9) stride 9: in place
    idx, num = idx.cpu().numpy(), num.cpu().numpy()
    assert np.array_equal(num, mask.sum(1))
    assert (num > cap).any() and (num < cap).any()
    for t in range(T):
        w = np.flatnonzero(mask[t])[:cap]
        assert np.array_equal(idx[t, :len(w)], w) and np.all(idx[t, len(w):] == -1), t
    # empty inputs
    i0, n0 = track.crop_points_in_boxes(p_d[:0], b_d, 8)
    assert torch.all(n0 == 0) and torch.all(i0 == -1)
