"""shared builders for the parity tests"""
import numpy as np
import torch

from detzero_b200.config import AttrDict

SMALL_RANGE = [-9.6, -9.6, -2, 9.6, 9.6, 4]
VOXEL = [0.1, 0.1, 0.15]
WAYMO_RANGE = [-75.2, -75.2, -2, 75.2, 75.2, 4]


def config1_cloud(n=20000, seed=0, pc_range=WAYMO_RANGE, c=5):
    """uniform cloud + 2 % out-of-range + 5 % exact duplicates / on-lattice coordinates"""
    g = np.random.default_rng(seed)
    lo, hi = np.array(pc_range[:3], np.float32), np.array(pc_range[3:], np.float32)
    p = g.uniform(lo, hi, (n, 3)).astype(np.float32)
    n_out = n // 50
    p[:n_out] = p[:n_out] * 1.2 + np.array([3.0, -3.0, 5.0], np.float32)          # outside (some in z only)
    n_dup = n // 20
    src = g.integers(n_out, n, n_dup)
    dst = g.integers(n_out, n, n_dup)
    p[dst[: n_dup // 2]] = p[src[: n_dup // 2]]                                     # exact duplicates
    k = g.integers(-700, 700, (n_dup - n_dup // 2, 3)).astype(np.float32)
    p[dst[n_dup // 2:]] = (k * np.array(VOXEL, np.float32)).astype(np.float32)      # exactly on the lattice
    p[g.integers(0, n, 8), 0] = hi[0]                                               # x == upper bound exactly
    feats = np.concatenate([np.tanh(g.uniform(0, 2, (n, 1))), g.uniform(0, 1, (n, max(c - 4, 1)))], axis=1)
    return np.concatenate([p, feats.astype(np.float32)], axis=1)[:, :c].astype(np.float32)


def clustered_cloud(n, seed, pc_range=SMALL_RANGE, c=5):
    """points concentrated on a few surfaces so that voxels have neighbours (exercises the rulebook)"""
    g = np.random.default_rng(seed)
    lo, hi = np.array(pc_range[:3], np.float32), np.array(pc_range[3:], np.float32)
    xy = g.uniform(lo[:2], hi[:2], (n, 2))
    z = np.where(g.random(n) < 0.6, g.normal(0.0, 0.05, n), g.uniform(0, 2.5, n))
    wall = g.random(n) < 0.3
    xy[wall, 0] = np.round(xy[wall, 0] / 3.0) * 3.0 + g.normal(0, 0.03, wall.sum())
    p = np.concatenate([xy, z[:, None]], axis=1)
    f = np.concatenate([np.tanh(g.uniform(0, 2, (n, 1))), g.uniform(0, 1, (n, c - 4))], axis=1)
    return np.concatenate([p, f], axis=1).astype(np.float32)


def model_cfg(backbone='VoxelResBackBone8x', mode='fp32', channels=None):
    cfg = AttrDict({
        'NAME': 'CenterPoint', 'SECOND_STAGE': False,
        'VFE': {'NAME': 'MeanVFE'},
        'BACKBONE_3D': {'NAME': backbone, 'COMPUTE_MODE': mode},
        'MAP_TO_BEV': {'NAME': 'HeightCompression', 'NUM_BEV_FEATURES': 256},
        'BACKBONE_2D': {'NAME': 'BaseBEVBackbone', 'LAYER_NUMS': [5, 5], 'LAYER_STRIDES': [1, 2],
                        'NUM_FILTERS': [128, 256], 'UPSAMPLE_STRIDES': [1, 2], 'NUM_UPSAMPLE_FILTERS': [256, 256],
                        'COMPUTE_MODE': mode},
        'DENSE_HEAD': {
            'NAME': 'CenterHead', 'CLASS_AGNOSTIC': False, 'COMPUTE_MODE': mode,
            'CLASS_NAMES_EACH_HEAD': [['Vehicle', 'Pedestrian', 'Cyclist']],
            'SHARED_CONV_CHANNEL': 64, 'USE_BIAS_BEFORE_NORM': True, 'NUM_HM_CONV': 2, 'IOU_WEIGHT': 1,
            'SEPARATE_HEAD_CFG': {
                'HEAD_ORDER': ['center', 'center_z', 'dim', 'rot', 'iou'],
                'HEAD_DICT': {'center': {'out_channels': 2, 'num_conv': 2}, 'center_z': {'out_channels': 1, 'num_conv': 2},
                              'dim': {'out_channels': 3, 'num_conv': 2}, 'rot': {'out_channels': 2, 'num_conv': 2},
                              'iou': {'out_channels': 1, 'num_conv': 2}}},
            'TARGET_ASSIGNER_CONFIG': {'FEATURE_MAP_STRIDE': 8, 'NUM_MAX_OBJS': 500, 'GAUSSIAN_OVERLAP': 0.1, 'MIN_RADIUS': 2},
            'POST_PROCESSING': {'SCORE_THRESH': 0.03, 'POST_CENTER_LIMIT_RANGE': [-80, -80, -10.0, 80, 80, 10.0],
                                'MAX_OBJ_PER_SAMPLE': 500,
                                'NMS_CONFIG': {'NMS_TYPE': 'nms_gpu', 'NMS_THRESH': 0.7, 'NMS_PRE_MAXSIZE': 4096,
                                               'NMS_POST_MAXSIZE': 500}}},
        'POST_PROCESSING': {'RECALL_THRESH_LIST': [0.3, 0.5, 0.7], 'SCORE_THRESH': 0.03, 'OUTPUT_RAW_SCORE': False,
                            'EVAL_METRIC': 'waymo'},
    })
    if channels is not None:
        cfg.BACKBONE_3D.CHANNELS = channels
    return cfg


CLASS_NAMES = ['Vehicle', 'Pedestrian', 'Cyclist']


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
