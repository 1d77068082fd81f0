"""GPU tests of the on-disk frame path: reference-format frame files -> pinned async reader -> device
preparation (dz_prepare_points) == the reference's host-side merge_sweeps + collate (oracle restatement, itself checked against
the reference's own source in tests/test_host_logic.py)."""
import numpy as np
import pytest
import torch

from oracle import det_ref

pytestmark = pytest.mark.gpu


def _pose(g):
    a = g.uniform(-0.3, 0.3)
    p = np.eye(4)
    p[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    p[:3, 3] = g.uniform(-5, 5, 3)
    return p


def _raw_frame(seed, n):
    from detzero_b200.det.dataset import synth_waymo_cloud
    g = np.random.default_rng(seed)
    p = synth_waymo_cloud(seed, n)                                     # x y z intensity(already tanh'ed) elongation offset
    raw = p.copy()
    raw[:, 3] = g.uniform(0, 3, p.shape[0])                            # raw intensity
    raw[:, 5] = np.where(g.random(p.shape[0]) < 0.93, -1.0, 1.0)       # NLZ flag
    return raw.astype(np.float32)


def test_pipeline_single_sweep_and_multi_sweep(cuda, tmp_path):
    from detzero_b200 import io as dzio
    g = np.random.default_rng(0)
    frames = []
    for i in range(6):
        raw = _raw_frame(10 + i, 20000 + 1000 * i)
        path = str(tmp_path / ('%04d.npy' % i))
        dzio.write_frame_npy(path, raw)
        frames.append({'path': path, 'raw': raw, 'pose': _pose(g), 'time_stamp': 1550000000000000 + 100000 * i, 'frame_id': '%04d' % i})
    # ---- 1-sweep: batches of 2 frames, identity transform, no time column
    batches = [[{'path': f['path'], 'frame_id': f['frame_id']} for f in frames[k:k + 2]] for k in (0, 2, 4)]
    pipe = dzio.FramePipeline(batches, cuda, max_points=30000)
    seen = 0
    for k, bd in enumerate(pipe):
        n = int(bd['points_count'].item())
        got = bd['points'][:n].cpu().numpy()
        want = []
        for b, f in enumerate(frames[2 * k:2 * k + 2]):
            info = {'pose': np.eye(4), 'time_stamp': 0}
            m = det_ref.merge_sweeps(info, [info], [f['raw'].copy()])[:, :5]
            want.append(np.concatenate([np.full((m.shape[0], 1), b), m], axis=1))
        want = np.concatenate(want).astype(np.float32)
        assert got.shape == want.shape
        assert np.array_equal(got[:, [0, 1, 2, 3, 5]], want[:, [0, 1, 2, 3, 5]])        # batch idx, xyz, elongation: bit-exact, file order
        assert np.abs(got[:, 4] - want[:, 4]).max() < 1e-6                               # tanh: libm vs CUDA
        seen += 1
    assert seen == 3 and pipe.bytes_read == sum(f['raw'].nbytes for f in frames)
    # ---- 3 sweeps per frame with ego-motion poses and the time column (waymo_5sweeps-style input)
    specs = []
    for cur in (2, 5):
        sw = [frames[cur - s] for s in range(3)]
        specs.append({'sweeps': [{'path': s['path'], 'pose': s['pose'], 'time_stamp': s['time_stamp']} for s in sw], 'pose': frames[cur]['pose'],
                      'time_stamp': frames[cur]['time_stamp'], 'frame_id': frames[cur]['frame_id'], '_sw': sw})
    pipe = dzio.FramePipeline([specs], cuda, max_points=30000, with_time=True)
    bd = next(iter(pipe))
    n = int(bd['points_count'].item())
    got = bd['points'][:n].cpu().numpy()
    want = []
    for b, sp in enumerate(specs):
        m = det_ref.merge_sweeps(sp['_sw'][0], sp['_sw'], [s['raw'].copy() for s in sp['_sw']])
        want.append(np.concatenate([np.full((m.shape[0], 1), b), m], axis=1))
    want = np.concatenate(want)
    assert got.shape == want.shape
    assert np.abs(got[:, 1:4] - want[:, 1:4].astype(np.float32)).max() < 2e-5          # double-precision transform, rounded to fp32 once
    assert np.array_equal(got[:, [0, 5]], want[:, [0, 5]].astype(np.float32))
    assert np.abs(got[:, 6] - want[:, 6]).max() < 1e-7 and np.abs(got[:, 4] - want[:, 4]).max() < 1e-6


def test_pipeline_feeds_the_detector(cuda, tmp_path):
    """files -> FramePipeline -> CenterPoint gives the detections of the in-memory path on the same points"""
    from tests import util
    from detzero_b200 import io as dzio, synthetic
    from detzero_b200.det import build_network
    from detzero_b200.det.dataset import SyntheticWaymoDataset, default_waymo_1sweep_cfg
    dcfg = default_waymo_1sweep_cfg()
    dcfg.POINT_CLOUD_RANGE = util.SMALL_RANGE
    ds = SyntheticWaymoDataset(dcfg, util.CLASS_NAMES, training=False, num_frames=2, n_points=30000)
    model = build_network(synthetic.model_cfg('VoxelBackBone8x', 'tf32'), 3, ds).eval()
    synthetic.load_seeded(model, 21)
    model = model.to(cuda)
    specs, mem = [], []
    for i in range(2):
        c = util.clustered_cloud(30000, 61 + i, c=5)
        raw = np.concatenate([c[:, :3], np.arctanh(np.clip(c[:, 3:4], 0, 0.999)), c[:, 4:5], -np.ones((c.shape[0], 1))], 1).astype(np.float32)
        path = str(tmp_path / ('f%d.npy' % i))
        dzio.write_frame_npy(path, raw)
        specs.append({'path': path, 'frame_id': str(i)})
        mem.append(raw)
    bd = next(iter(dzio.FramePipeline([specs], cuda, max_points=40000)))
    n = int(bd['points_count'].item())
    pts = bd['points'][:n]
    sizes = torch.bincount(pts[:, 0].long(), minlength=2).tolist()
    with torch.no_grad():
        a = model({'points': pts.contiguous(), 'points_per_frame': sizes, 'batch_size': 2, 'frame_id': bd['frame_id']})[0]
        host = np.concatenate([np.concatenate([np.full((m.shape[0], 1), b, np.float32), m[:, :3], np.tanh(m[:, 3:4]), m[:, 4:5]], 1) for b, m in enumerate(mem)])
        keep = np.concatenate([(np.abs(m[:, 0]) <= 9.6) & (np.abs(m[:, 1]) <= 9.6) for m in mem])
        b = model({'points': torch.from_numpy(host.astype(np.float32)).to(cuda), 'points_per_frame': [m.shape[0] for m in mem], 'batch_size': 2,
                   'frame_id': bd['frame_id']})[0]
    from tests.test_gpu_det import _assert_same_detections
    for k in range(2):                              # same detections up to the 1-ulp tanh difference of the intensity feature
        _assert_same_detections(a[k], {kk: v.cpu() for kk, v in b[k].items()}, score_tol=1e-3, box_tol=5e-3, count_slack=2)
