"""world_size-2 gloo test (CPU) of the N>1 host logic: frame sharding order and the per-sequence box gather reproduce
the reference's merge_results_dist ordering (common_utils.py:135-138)."""
import json
import os
import socket
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_indices_match_reference_sampler():
    from detzero_b200.dist import shard_indices
    assert shard_indices(7, 0, 2) == [0, 2, 4, 6] and shard_indices(7, 1, 2) == [1, 3, 5, 0]      # tail wraps
    assert shard_indices(199, 3, 8)[:3] == [3, 11, 19] and len(shard_indices(199, 3, 8)) == 25
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]


def test_gather_world2(tmp_path):
    port, num_frames, out = _free_port(), 7, str(tmp_path / 'out.json')
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, '_dist_worker.py'), str(r), '2', str(port), str(num_frames), out])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    res = json.load(open(out))
    assert res['tags'] == [float(f) for f in range(num_frames)]          # frame order restored, truncated to length
    assert res['counts'] == [f % 4 + 1 for f in range(num_frames)]
    assert res['tags2'] == res['tags'] and res['counts2'] == res['counts']      # SequenceGather: same order from ONE collective
