"""Deterministic, construction-order-independent weights and inputs for parity tests (TEST INFRASTRUCTURE).

There are no released checkpoints on disk, so both the reference modules (when generating golden
vectors) and the product modules are filled from the SAME seeded function of (key name, shape): identical key names
=> identical tensors, regardless of how each side constructs its module tree.  BatchNorm running stats are
randomised away from (0, 1) so that folding errors show."""
import numpy as np

from detzero_b200.synthetic import _rng, seeded_tensor, seeded_state_dict, load_seeded    # noqa: F401  (one definition, shared)


def random_sparse_coords(seed, B, shape, density):
    """unique (b,z,y,x) int32 sites in random order"""
    g = np.random.default_rng(seed)
    dense = g.random((B, *shape)) < density
    idx = np.argwhere(dense).astype(np.int32)
    return idx[g.permutation(len(idx))]
