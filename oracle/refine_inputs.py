"""Seeded refiner inputs + the REGRESSION sections of the shipped refiner configs
(refining/tools/cfgs/ref_model_cfgs/vehicle_{prm,grm,crm}_model.yaml) -- TEST INFRASTRUCTURE."""
import numpy as np
import torch

from detzero_b200.config import AttrDict


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
