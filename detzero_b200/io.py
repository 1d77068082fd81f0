"""On-disk formats of the reference + a pinned-memory, asynchronous frame reader.

Formats (kept byte-compatible so real Waymo data prepared by the reference's tools feeds the kernels as is):
  * per-frame lidar file ``<frame>.npy``: (N, 6) float32 ``[x, y, z, intensity, elongation, NLZ_flag]``
    (detection/detzero_det/datasets/waymo/waymo_utils.py:284-302, read at waymo_dataset.py:93-102)
  * detection results ``result.pkl``: list of per-frame dicts ``name / score / boxes_lidar / sequence_name / frame_id / pose``
    (detection/detzero_det/datasets/dataset.py:306-354); ``sequence_list_to_dict`` regroups them per sequence for the tracker
    (tracking/detzero_track/utils/data_utils.py:15-22).

Reader: the file bytes go disk -> pinned host buffer (``readinto``, worker threads, no intermediate copy) -> device
(``cudaMemcpyAsync`` on a copy stream) untouched; what the reference does on the host per frame -- NLZ filter, tanh(intensity),
ego-motion transform, time column, collate batch column (dataset.py:167-196,275-283) -- runs on the device
(``dz_prepare_points``, csrc/io.cu).  While batch k is computed, the files of batch k+1 are read and copied.
"""
import ast
import concurrent.futures as cf
import pickle
import struct

import numpy as np
import torch

from . import ops
from ._lib import check, lib


# ---- .npy ------------------------------------------------------------------------------------------------------------------
def read_npy_header(f):
    """-> (shape, dtype, fortran_order, data_offset) of an open .npy file (format 1.0 / 2.0 / 3.0)"""
    magic = f.read(6)
    if magic != b'\x93NUMPY':
        raise ValueError('not a .npy file')
    major, _minor = f.read(2)
    if major == 1:
        hlen = struct.unpack('<H', f.read(2))[0]
    else:
        hlen = struct.unpack('<I', f.read(4))[0]
    d = ast.literal_eval(f.read(hlen).decode('latin1'))
    return tuple(d['shape']), np.dtype(d['descr']), bool(d['fortran_order']), f.tell()


def read_frame_into(path, pinned):
    """read a (N, 6) float32 frame file straight into a pinned (cap, 6) float32 tensor; returns N"""
    with open(path, 'rb') as f:
        shape, dtype, fortran, _ = read_npy_header(f)
        if len(shape) != 2 or shape[1] != 6 or dtype != np.dtype('<f4') or fortran:
            raise ValueError('%s: expected a C-ordered (N, 6) float32 frame file, got %s %s' % (path, shape, dtype))
        n = shape[0]
        if n > pinned.shape[0]:
            raise ValueError('%s: %d points > buffer capacity %d' % (path, n, pinned.shape[0]))
        view = memoryview(pinned.numpy()).cast('B')[:n * 24]
        got = f.readinto(view)
        if got != n * 24:
            raise IOError('%s: truncated (%d of %d bytes)' % (path, got, n * 24))
    return n


def write_frame_npy(path, points6):
    """the reference's frame file: (N, 6) float32 [x, y, z, intensity, elongation, NLZ_flag] (waymo_utils.py:296-300)"""
    a = np.ascontiguousarray(points6, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 6
    np.save(path, a)


# ---- result.pkl ------------------------------------------------------------------------------------------------------------
def generate_prediction_dicts(batch_dict, pred_dicts, class_names):
    """dataset.py:306-354: per-frame annos from the detector's pred_dicts"""
    annos = []
    for index, box_dict in enumerate(pred_dicts):
        scores = box_dict['pred_scores'].cpu().numpy()
        boxes = box_dict['pred_boxes'].cpu().numpy()
        labels = box_dict['pred_labels'].cpu().numpy()
        d = {'name': np.zeros(0), 'score': np.zeros(0), 'boxes_lidar': np.zeros([0, 9])}
        if scores.shape[0]:
            d = {'name': np.array(class_names)[labels - 1], 'score': scores, 'boxes_lidar': boxes}
        for k in ('sequence_name', 'frame_id', 'pose'):
            if k in batch_dict:
                d[k] = batch_dict[k][index]
        annos.append(d)
    return annos


def save_result_pkl(path, annos):
    with open(path, 'wb') as f:
        pickle.dump(annos, f)


def load_result_pkl(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


def sequence_list_to_dict(annos):
    """tracking/detzero_track/utils/data_utils.py:15-22: {sequence_name: {frame_id: anno}}"""
    out = {}
    for a in annos:
        out.setdefault(a['sequence_name'], {})[a['frame_id']] = a
    return out


def gathered_to_annos(boxes, counts, class_names, frame_ids=None, sequence_name='sequence', poses=None):
    """the all-gathered device tensor (F, 500, 9) + counts (F,) -> the reference's per-frame anno dicts (ONE D2H of each)"""
    b, c = boxes.cpu().numpy(), counts.cpu().numpy()
    annos = []
    for f in range(b.shape[0]):
        rows = b[f, :c[f]]
        annos.append({'name': np.array(class_names)[rows[:, 8].astype(np.int64) - 1] if len(rows) else np.zeros(0),
                      'score': rows[:, 7].copy(), 'boxes_lidar': rows[:, :7].copy(), 'sequence_name': sequence_name,
                      'frame_id': frame_ids[f] if frame_ids is not None else '%04d' % f, 'pose': poses[f] if poses is not None else np.eye(4)})
    return annos


# ---- device preparation -------------------------------------------------------------------------------------------------------
def prepare_points(raw, n, out, d_count, batch_idx, transform=None, time_offset=0.0, with_time=False):
    """append the prepared points of one raw frame / sweep (device (>=n, 6) f32) to ``out`` at ``d_count[0]``; see dz_prepare_points"""
    assert raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous() and raw.shape[1] == 6
    C = 1 + 5 + (1 if with_time else 0)
    assert out.shape[1] == C and out.is_contiguous()
    t12 = None
    if transform is not None:
        t = np.ascontiguousarray(np.asarray(transform, dtype=np.float64)[:3, :4])
        t12 = (ctypes_double * 12)(*t.reshape(-1).tolist())
    ws = ops.workspace(lib().dz_prepare_points_ws_bytes(int(n)), raw.device, 'prep')
    check(lib().dz_prepare_points(ops._p(raw), int(n), t12, float(time_offset), int(with_time), int(batch_idx), ops._p(out), out.shape[0],
                                  ops._p(d_count), ops._p(ws), ws.numel(), ops._stream()), 'prepare_points')
    ops._count(3)


import ctypes as _ct  # noqa: E402
ctypes_double = _ct.c_double


class FramePipeline:
    """Asynchronous reader: ``for batch in FramePipeline(batches_of_frame_specs, ...)`` yields collated device batches
    ``{'points': (N, 1+C) f32, 'voxel_count' ... }`` ready for CenterPoint.forward_device while the next batch's files are being read
    (thread pool, pinned buffers) and copied (copy stream).

    A frame spec is ``{'path': str}`` or, for multi-sweep input (dataset.py:140-196), ``{'sweeps': [{'path', 'pose', 'time_stamp'}, ...],
    'pose': current_pose, 'time_stamp': current_time}`` (first sweep = the current frame)."""

    def __init__(self, batches, device, max_points=250000, with_time=False, workers=8, depth=2):
        self.batches, self.dev, self.max_points, self.with_time = batches, torch.device(device), int(max_points), bool(with_time)
        self.pool = cf.ThreadPoolExecutor(max_workers=workers)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        n_sw = max(len(f.get('sweeps', [f])) for b in batches for f in b)
        per_batch = max(len(b) for b in batches) * n_sw
        self.slots = [{'pinned': [torch.empty((self.max_points, 6), dtype=torch.float32).pin_memory() for _ in range(per_batch)],
                       'dev': [torch.empty((self.max_points, 6), dtype=torch.float32, device=self.dev) for _ in range(per_batch)],
                       'event': torch.cuda.Event(), 'free': torch.cuda.Event()} for _ in range(depth)]
        self.depth = depth
        self.bytes_read = 0

    @staticmethod
    def _sweeps(spec):
        return spec['sweeps'] if 'sweeps' in spec else [{'path': spec['path'], 'pose': None, 'time_stamp': 0}]

    def _load(self, k, slot):
        """disk -> pinned (threads) -> device (copy stream); returns the per-sweep point counts"""
        files = [s['path'] for f in self.batches[k] for s in self._sweeps(f)]
        slot['free'].synchronize()                     # the consumer of this slot's previous contents has finished (host-side wait)
        counts = list(self.pool.map(lambda a: read_frame_into(*a), [(p, slot['pinned'][i]) for i, p in enumerate(files)]))
        with torch.cuda.stream(self.copy_stream):
            for i, n in enumerate(counts):
                slot['dev'][i][:n].copy_(slot['pinned'][i][:n], non_blocking=True)
            slot['event'].record(self.copy_stream)
        self.bytes_read += sum(counts) * 24
        return counts

    def __iter__(self):
        nb = len(self.batches)
        futures = {}
        loader = cf.ThreadPoolExecutor(max_workers=1)
        for k in range(min(self.depth - 1, nb)):
            futures[k] = loader.submit(self._load, k, self.slots[k % self.depth])
        for k in range(nb):
            nxt = k + self.depth - 1
            if nxt < nb:
                futures[nxt] = loader.submit(self._load, nxt, self.slots[nxt % self.depth])
            slot = self.slots[k % self.depth]
            counts = futures.pop(k).result()
            torch.cuda.current_stream().wait_event(slot['event'])
            yield self._collate(self.batches[k], slot, counts)
            slot['free'].record(torch.cuda.current_stream())
        loader.shutdown()

    def _collate(self, batch, slot, counts):
        C = 1 + 5 + (1 if self.with_time else 0)
        total = sum(counts)
        out = torch.empty((max(total, 1), C), dtype=torch.float32, device=self.dev)
        d_count = torch.zeros(2, dtype=torch.int32, device=self.dev)
        i = 0
        for b, spec in enumerate(batch):
            cur_pose = spec.get('pose')
            for s in self._sweeps(spec):
                T, dt = None, 0.0
                if s.get('pose') is not None and cur_pose is not None:
                    T = np.linalg.inv(np.asarray(cur_pose, np.float64)) @ np.asarray(s['pose'], np.float64)      # dataset.py:186
                    dt = float(int(s['time_stamp']) - int(spec['time_stamp'])) / 1000000.0                       # :187,190
                prepare_points(slot['dev'][i], counts[i], out, d_count, b, transform=T, time_offset=dt, with_time=self.with_time)
                i += 1
        return {'points': out, 'points_count': d_count[0:1], 'batch_size': len(batch), 'raw_counts': counts,
                'frame_id': np.array([str(f.get('frame_id', k)) for k, f in enumerate(batch)])}
