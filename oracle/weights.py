"""Deterministic, construction-order-independent weights and inputs for parity tests (TEST INFRASTRUCTURE).

There are no released checkpoints on disk, so both the reference modules (when generating golden
vectors) and the product modules are filled from the SAME seeded function of (key name, shape): identical key names
=> identical tensors, regardless of how each side constructs its module tree.  BatchNorm running stats are
randomised away from (0, 1) so that folding errors show."""
import zlib

import numpy as np
import torch


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


def random_sparse_coords(seed, B, shape, density):
    """unique (b,z,y,x) int32 sites in random order"""
    g = np.random.default_rng(seed)
    dense = g.random((B, *shape)) < density
    idx = np.argwhere(dense).astype(np.int32)
    return idx[g.permutation(len(idx))]
