"""Device-side pieces of the steps around the detector / refiner, with the reference's function
names so that the tracker / daemon can import them in place of the originals:

  * ``IoUBEV_dis_mat`` / ``IoU3D_dis_mat`` / ``IoU2D_dis_mat`` / ``bev_overlap_gpu`` --
    tracking/detzero_track/models/tracking_modules/data_association/distance.py:44-141.  They take CUDA tensors -- e.g. rows of
    the all-gathered ``(F, 500, 9)`` detection tensor, no ``result.pkl`` hop -- and return CUDA tensors (the reference copies every
    matrix to the host; the Hungarian assignment itself stays on the host and is out of scope).
  * ``points_in_boxes_gpu_v2`` / ``crop_points_in_boxes`` -- utils/detzero_utils/ops/roiaware_pool3d/roiaware_pool3d_utils.py:45-58
    and the per-object crop of daemon/prepare_object_data.py:264-311.
"""
import ctypes

import torch

from . import ops
from ._lib import check, lib

IOU_BEV, OVERLAP_BEV, IOU_3D, IOU_2D = 0, 1, 2, 3


def _rows7(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] >= 7 and t.stride(1) == 1, (t.shape, t.dtype)
    return t


def boxes_pairwise(boxes_a, boxes_b, kind):
    """(N, >=7), (M, >=7) CUDA float rows (any row stride) -> (N, M) CUDA matrix"""
    a, b = _rows7(boxes_a), _rows7(boxes_b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    if a.shape[0] and b.shape[0]:
        check(lib().dz_boxes_pairwise(ops._p(a), a.shape[0], a.stride(0), ops._p(b), b.shape[0], b.stride(0), int(kind), ops._p(out),
                                      ops._stream()), 'boxes_pairwise')
        ops._count(1)
    return out


def IoUBEV_dis_mat(boxes_a, boxes_b, gpu=True):
    return boxes_pairwise(boxes_a, boxes_b, IOU_BEV)


def IoU3D_dis_mat(boxes_a, boxes_b):
    return boxes_pairwise(boxes_a, boxes_b, IOU_3D)


def bev_overlap_gpu(boxes_a, boxes_b):
    return boxes_pairwise(boxes_a, boxes_b, OVERLAP_BEV)


def IoU2D_dis_mat(boxes_a, boxes_b):
    """(N,4), (M,4) [x, y, w, h] like the reference"""
    def pad(t):
        o = torch.zeros((t.shape[0], 7), dtype=torch.float32, device=t.device)
        o[:, 0:2], o[:, 3:5] = t[:, 0:2], t[:, 2:4]
        return o
    return boxes_pairwise(pad(boxes_a), pad(boxes_b), IOU_2D)


dis_mat_dict = {'IoU2D': IoU2D_dis_mat, 'IoU3D': IoU3D_dis_mat, 'IoUBEV': IoUBEV_dis_mat}     # data_association/__init__.py:8-13


def points_in_boxes_gpu_v2(points, boxes):
    """points (B, M, 3), boxes (B, T, 7) -> (B, T, M) int32 mask (the reference's output form)"""
    assert points.shape[0] == boxes.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    out = torch.zeros((points.shape[0], boxes.shape[1], points.shape[1]), dtype=torch.int32, device=points.device)
    for b in range(points.shape[0]):
        p, bx = points[b].contiguous().float(), boxes[b].contiguous().float()
        check(lib().dz_points_in_boxes_mask(ops._p(p), p.shape[0], 3, ops._p(bx), bx.shape[0], 7, ops._p(out[b]), ops._stream()),
              'points_in_boxes_mask')
        ops._count(1)
    return out


def crop_points_in_boxes(points, boxes, cap):
    """points (M, C>=3) [x,y,z,...], boxes (T, >=7) -> idx (T, cap) int32 (indices of each box's points in input order, -1 padded),
    num (T,) int32 true counts.  No (T, M) mask is materialised, nothing goes to the host."""
    assert points.is_cuda and points.dtype == torch.float32 and points.stride(1) == 1
    bx = _rows7(boxes)
    T, M = bx.shape[0], points.shape[0]
    idx = torch.empty((T, cap), dtype=torch.int32, device=points.device)
    num = torch.zeros((T,), dtype=torch.int32, device=points.device)
    if M == 0:
        idx.fill_(-1)
    elif T:
        ws = ops.workspace(lib().dz_crop_points_ws_bytes(M, T), points.device, 'crop')
        check(lib().dz_crop_points_in_boxes(ops._p(points), M, points.stride(0), ops._p(bx), T, bx.stride(0), ops._p(idx), int(cap),
                                            ops._p(num), ops._p(ws), ws.numel(), ops._stream()), 'crop_points_in_boxes')
        ops._count(3)
    return idx, num
