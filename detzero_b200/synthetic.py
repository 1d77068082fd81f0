"""Seeded synthetic weights and the default CenterPoint model config for benchmarks, smoke runs and parity tests.

There are no released checkpoints on disk: product modules, the oracle and the reference modules (when
golden vectors are generated) are all filled from the SAME seeded function of (key name, shape) -- identical key names
=> identical tensors, regardless of how each side constructs its module tree.  BatchNorm running stats are randomised
away from (0, 1) so that folding errors show.  Lives in the package so that bench.py's GPU arm does not need ``oracle/``
or ``tests/`` on its import path."""
import zlib

import numpy as np
import torch

from .config import AttrDict

#: sparse-conv arithmetic of the shipped config / bench / smoke (detzero_b200/_lib.py MODES), and the per-mode bound on the
#: backbone output features vs the fp32 oracle (relative to max |feature|) that the parity tests enforce
DEFAULT_SP_MODE = 'bf16x2'
SP_MODE_TOL = {'fp32': 2e-5, 'tf32x3': 2e-4, 'bf16x2': 2e-4, 'tf32': 5e-3, 'bf16': 3e-2}

def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode())])


def seeded_tensor(key, shape, seed, kind=None):
    shape = tuple(int(s) for s in shape)
    g = _rng(seed, key)
    leaf = key.split('.')[-1]
    if '.hm.' in key and leaf == 'bias' and len(shape) == 1 and kind is None and key.endswith('.1.bias'):
        return torch.from_numpy((-2.19 + g.normal(0, 0.05, shape)).astype(np.float32))     # center_head.py:33 init
    if '.hm.' in key and key.endswith('.1.weight'):
        fan_in = max(1, int(np.prod(shape)) // max(1, shape[0]))
        return torch.from_numpy((g.normal(0, 0.15 * np.sqrt(2.0 / fan_in), shape)).astype(np.float32))
    if leaf == 'num_batches_tracked':
        return torch.zeros(shape, dtype=torch.long)
    if leaf == 'running_var':
        a = g.uniform(0.5, 1.5, shape)
    elif leaf == 'running_mean':
        a = g.normal(0, 0.1, shape)
    elif leaf == 'bias' or leaf == 'in_proj_bias':
        a = g.normal(0, 0.05, shape)
    elif leaf == 'weight' and len(shape) == 1:           # norm scale
        a = g.uniform(0.5, 1.5, shape)
    else:                                                # conv / linear weight: He-style, fan_in = prod(shape[1:])...
        fan_in = max(1, int(np.prod(shape)) // max(1, shape[0]))
        a = g.normal(0, np.sqrt(2.0 / fan_in), shape)
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))


def seeded_state_dict(module, seed):
    """state_dict with every entry replaced by its seeded value (same dtype/shape)"""
    sd = module.state_dict()
    out = {}
    for k in sd:
        t = seeded_tensor(k, sd[k].shape, seed)
        out[k] = t.to(sd[k].dtype) if sd[k].dtype != torch.long else t
    return out


def load_seeded(module, seed):
    sd = seeded_state_dict(module, seed)
    module.load_state_dict(sd)
    return sd



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


# ---- refiner: the REGRESSION sections of the shipped configs (refining/tools/cfgs/ref_model_cfgs/vehicle_{prm,grm,crm}_model.yaml)
# ---- and seeded inputs of the reference's shapes
def prm_cfg():
    return AttrDict({'NAME': 'PositionTransformer', 'QUERY_ENCODER': [128, 128], 'MEMORY_ENCODER': [128, 128], 'REGRESSION_MLP': [512],
                     'DECODER': {'NAME': 'PositionHead', 'num_classes': 3, 'num_heads': 8, 'num_decoder_layers': 1, 'auxiliary': True,
                                 'cross_only': False, 'hidden_channel': 256, 'dropout': 0.1, 'bn_momentum': 0.1, 'activation': 'relu',
                                 'ffn_channel': 256},
                     'LOSS_CLS': {'type': 'CrossEntropyLoss', 'reduction': 'mean', 'ignore_index': -1}})


def grm_cfg():
    return AttrDict({'NAME': 'GeometryTransformer', 'QUERY_ENCODER': [128, 128], 'MEMORY_ENCODER': [128, 128], 'REGRESSION_MLP': [512],
                     'EMBED_DIMS': 256, 'ANCHOR_SIZES': [[4.8, 1.8, 1.5], [10.0, 2.6, 3.2], [2.0, 1.0, 1.6]],
                     'DECODER': {'NAME': 'GeometryHead', 'num_classes': 3, 'num_heads': 8, 'num_decoder_layers': 1, 'auxiliary': True,
                                 'cross_only': False, 'memory_self_attn': False, 'hidden_channel': 256, 'ffn_channel': 256,
                                 'dropout': 0.1, 'bn_momentum': 0.1, 'activation': 'relu'}})


def crm_cfg():
    return AttrDict({'NAME': 'ConfidencePointnet', 'ENCODER_MLP': [128, 128], 'REGRESSION_MLP': [512], 'SCORE_THRESH': [0.35, 0.7]})


def prm_inputs(seed, B=2, boxes=200, qpts=256, mpts=48, dims=32):
    g = np.random.default_rng(seed)
    box_num = g.integers(20, boxes + 1, B)
    box_num[0] = boxes                                   # one full track
    mask = np.zeros((B, boxes), np.float32)
    for b in range(B):
        mask[b, box_num[b]:] = 1
    return {'pos_query_points': torch.from_numpy(g.normal(0, 1, (B, boxes, qpts, dims)).astype(np.float32)),
            'pos_memory_points': torch.from_numpy(g.normal(0, 1, (B, boxes, mpts, dims)).astype(np.float32)),
            'pos_trajectory': torch.from_numpy(g.normal(0, 2, (B, boxes, 7)).astype(np.float32)),
            'padding_mask': torch.from_numpy(mask)}


def grm_inputs(seed, B=2, mem=4096, q=3, qpts=256):
    g = np.random.default_rng(seed)
    return {'geo_memory_points': torch.from_numpy(g.normal(0, 1, (B, mem, 11)).astype(np.float32)),
            'geo_query_points': torch.from_numpy(g.normal(0, 1, (B, q, qpts, 4)).astype(np.float32)),
            'geo_query_boxes': torch.from_numpy(g.normal(0, 1, (B, q, 7)).astype(np.float32)),
            'geo_query_num': torch.tensor([q] + [max(1, q - 1)] * (B - 1))}


def crm_inputs(seed, B=2, boxes=200, pts=256, dims=32):
    g = np.random.default_rng(seed)
    return {'conf_points': torch.from_numpy(g.normal(0, 1, (B, boxes, pts, dims)).astype(np.float32))}
