#!/usr/bin/env python
"""bench.py -- Waymo-shape frames/sec through the B200-native DetZero detector hot path.

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic input: raw points -> hard voxelization (+MeanVFE)
-> rulebooks -> sparse VoxelBackBone8x -> BEV scatter -> BEV backbone -> CenterHead -> decode -> rotated NMS.
Workload (BASELINE.json configs[1]): CenterPoint 1-sweep, VoxelBackBone8x, synthetic 180 K-pt Waymo-range cloud,
fp32 storage, 8 frames per step per GPU by default (the reference's own BATCH_SIZE_PER_GPU, centerpoint_1sweep.yaml:88;
`--batch 1` gives the single-frame latency configuration).  Frames shard across ranks with no data-path collective (weak
scaling); see detzero_b200/dist.py for the per-sequence box gather that is NOT part of a step.

Timing: CUDA events around every step on the launching stream, L2 flushed (256 MiB write) between steps outside the
timed intervals, max over ranks of the summed step times.  `value` has the input resident in HBM; `e2e` includes the
pinned-host -> device copy of the step's points and the device -> host read of the boxes, through the public
CenterPoint.forward(batch_dict) API.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POINTS = 180000
NUM_CLOUDS = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--backbone', default='VoxelBackBone8x', choices=['VoxelBackBone8x', 'VoxelResBackBone8x'])
    ap.add_argument('--mode', default=os.environ.get('DZ_MODE', 'tf32'),
                    help='dense BEV/head convs: tf32 (tcgen05; what the reference gets from cuDNN by default, SURVEY A.6) | fp32 (exact FMA)')
    ap.add_argument('--sp-mode', default=os.environ.get('DZ_SP_MODE', 'tf32'),
                    help='sparse-conv arithmetic: tf32 (tcgen05, 1 pass) | tf32x3 (tcgen05, hi/lo split) | fp32 (exact FMA, spconv default)')
    ap.add_argument('--batch', type=int, default=8,
                    help='frames per step and GPU; 8 = the reference config (centerpoint_1sweep.yaml:88 BATCH_SIZE_PER_GPU)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
    ap.add_argument('--no-schedule', action='store_true', help='run the tensor-core sparse convs without the mask-grouped tile schedule')
    ap.add_argument('--layer-times', action='store_true', help='print per-layer sparse-conv times of the traced frame to stderr')
    ap.add_argument('--modules', type=int, default=0, help='diagnostic: graph-replay time of only the first N modules of the detector (stderr, then exit)')
    ap.add_argument('--stage-times', action='store_true', help='print a per-stage device time table to stderr')
    return ap.parse_args()


def make_model_cfg(backbone, mode, sp_mode):
    from detzero_b200 import synthetic
    cfg = synthetic.model_cfg(backbone, mode)
    cfg.BACKBONE_3D.COMPUTE_MODE = sp_mode
    if os.environ.get('DZ_NO_OVERLAP'):
        cfg.BACKBONE_3D.OVERLAP_RULEBOOKS = False      # diagnostic: rulebooks inline on the main stream
    return cfg


def tune_head_for_bench(model):
    """The seeded random weights give a flat heatmap and a negative IoU map, i.e. no detections and an idle post-processing
    stage.  Re-scale the heatmap / IoU head outputs (same rule for the GPU arm and the CPU arm) so that a frame yields
    ~1.5 K candidates above SCORE_THRESH, 500 boxes into the rotated NMS (calibrated with the oracle on frame 0)."""
    import torch
    sd = model.state_dict()
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        sd['dense_head.heads_list.0.hm.0.1.bias'].copy_(torch.randn(64, generator=g) * 0.2)
        sd['dense_head.heads_list.0.hm.1.weight'].mul_(torch.tensor([340.865, 293.364, 329.442]).view(3, 1, 1, 1))
        sd['dense_head.heads_list.0.hm.1.bias'].copy_(torch.tensor([-7.4753, -0.0872, 14.5139]))
        sd['dense_head.heads_list.0.iou.1.bias'].fill_(0.85)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def build_inputs(batch):
    from detzero_b200 import synthetic
    from detzero_b200.det.dataset import SyntheticWaymoDataset, default_waymo_1sweep_cfg
    ds = SyntheticWaymoDataset(default_waymo_1sweep_cfg(), synthetic.CLASS_NAMES, training=False, num_frames=NUM_CLOUDS * batch,
                               n_points=N_POINTS)
    batches = []
    for i in range(NUM_CLOUDS):
        items = [ds[i * batch + j] for j in range(batch)]
        for it in items:                    # fixed shape for CUDA-graph replay: pad with far out-of-range points (dropped by
            p = it['points']                # the voxelizer exactly like any other out-of-range point)
            if p.shape[0] < N_POINTS:
                pad = np.full((N_POINTS - p.shape[0], p.shape[1]), 1.0e4, dtype=np.float32)
                it['points'] = np.concatenate([p, pad], axis=0)
        batches.append(ds.collate_batch(items))
    return ds, batches


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == 'active':
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path (voxelizer C restatement + spconv 'Native' + torch-CPU convs)
# ------------------------------------------------------------------------------------------------------------------
def cpu_frame_fn(backbone, seed=3):
    import torch
    import oracle
    from oracle import det_ref, weights
    from tests import util
    from detzero_b200.det import build_network
    ds, batches = build_inputs(1)
    model = build_network(make_model_cfg(backbone, 'fp32', 'fp32'), 3, ds).eval()
    weights.load_seeded(model, seed)
    sd = tune_head_for_bench(model)
    vox = oracle.Point2VoxelCPU3d(util.VOXEL, util.WAYMO_RANGE, 5, 5, 200000)
    post = dict(MAX_OBJ_PER_SAMPLE=500, SCORE_THRESH=0.03, POST_CENTER_LIMIT_RANGE=[-80, -80, -10.0, 80, 80, 10.0],
                NMS_THRESH=0.7, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)
    res = backbone == 'VoxelResBackBone8x'

    def run(i):
        pts = batches[i % len(batches)]['points'][:, 1:]
        with torch.no_grad():
            v, c, n = vox.point_to_voxel(pts)
            lv = det_ref.voxel_backbone(sd, 'backbone3d.', det_ref.mean_vfe(v, n), np.pad(c, ((0, 0), (1, 0))), [41, 1504, 1504], 1, res)
            s2d = det_ref.bev_backbone(sd, 'backbone2d.', det_ref.height_compression(lv['out']), [5, 5], [1, 2], [1, 2])
            maps = det_ref.center_head_maps(sd, 'dense_head.', s2d, ['center', 'center_z', 'dim', 'rot', 'iou', 'hm'])
            return det_ref.generate_predicted_boxes(maps, util.WAYMO_RANGE, util.VOXEL, 8, post, use_iou=True)
    return run


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (oracle port; spconv itself cannot be installed
    here).  One step = one frame.  Under torchrun only rank 0 works."""
    import torch
    if int(os.environ.get('RANK', '0')) != 0:
        return
    cores = cpu_threads()
    torch.set_num_threads(cores)
    run = cpu_frame_fn(args.backbone)
    for i in range(args.warmup):
        run(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {'impl': 'reference', 'metric': 'Waymo-shape frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * dt / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': 'CenterPoint 1-sweep %s, synthetic 180K-pt Waymo-range cloud, fp32, batch %d/GPU'
                                   % (args.backbone, args.batch),
                       'arm': 'oracle port of the reference algorithm on the host CPU (spconv cannot be installed here)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d frames of the same workload, torch threads=%d' % (args.steps, cores)},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from detzero_b200 import ops, synthetic as weights
    from detzero_b200.det import build_network

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    ds, batches = build_inputs(args.batch)
    if os.environ.get('DZ_SIDE_PRIO') == '0':
        from detzero_b200.det import backbone3d as _b3
        _b3._Backbone8xBase.SIDE_STREAM_HIGH_PRIORITY = False
    if args.no_schedule:
        from detzero_b200.spconv import pytorch as _sp
        _sp._SparseConv.SCHEDULE_TILES = False
    model = build_network(make_model_cfg(args.backbone, args.mode, args.sp_mode), 3, ds).eval()
    weights.load_seeded(model, 3)
    tune_head_for_bench(model)
    model = model.to(dev)

    # pinned host copies (e2e arm) and resident device copies (value arm); each rank takes its own frames
    host_pts = [torch.from_numpy(b['points']).pin_memory() for b in batches]
    dev_pts = [h.to(dev) for h in host_pts]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out_host = torch.empty((args.batch, 500, 9), dtype=torch.float32).pin_memory()

    def batch_dict(i, pts):
        b = batches[i % NUM_CLOUDS]
        return {'points': pts, 'points_per_frame': b['points_per_frame'], 'frame_id': b['frame_id'], 'batch_size': b['batch_size']}

    graph_state = {}
    e2e_state = {'d2h': 0, 'boxes': 0}

    def step_resident(i):
        if graph_state:
            graph_state['in'].copy_(dev_pts[i % NUM_CLOUDS], non_blocking=True)      # device->device, part of the step
            graph_state['g'].replay()
            return graph_state['out']
        with torch.no_grad():
            bd = model.forward_device(batch_dict(i, dev_pts[i % NUM_CLOUDS]))
        return bd

    def step_e2e(i):
        with torch.no_grad():
            if graph_state:
                graph_state['in'].copy_(host_pts[i % NUM_CLOUDS], non_blocking=True)     # pinned host -> device
                graph_state['g'].replay()
                pred, _ = model.post_processing(graph_state['out'])                      # the step's D2H (counts) + dicts
            else:
                pts = host_pts[i % NUM_CLOUDS].to(dev, non_blocking=True)
                pred, _ = model(batch_dict(i, pts))                   # public API: includes the D2H read of counts
            d2h = 4 * len(pred)                                        # the count read inside post_processing
            for b, pd in enumerate(pred):                              # every frame's boxes go back to pinned host memory
                n = pd['pred_boxes'].shape[0]
                out_host[b, :n, :7].copy_(pd['pred_boxes'], non_blocking=True)
                d2h += n * 7 * 4
            e2e_state['d2h'], e2e_state['boxes'] = d2h, sum(pd['pred_boxes'].shape[0] for pd in pred)
        return pred

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        ops.reset_launch_count()
        for i in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(warmup + i)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        launches = ops.launch_count() if not graph_state else graph_state.get('launches_per_step', 0) * steps
        if world > 1:
            dist.barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), launches

    # capacity hints (buffer sizes / launch grids follow the observed sparsity): a few full frames through the public API
    for rep in range(2):
        for i in range(NUM_CLOUDS):
            step_e2e(i)
    torch.cuda.synchronize()

    if args.stage_times and rank == 0:
        # main-stream device time per module of the eager frame (rulebook side streams overlap the backbone's convs)
        names = [type(m).__name__ for m in model.module_list]
        acc = [0.0] * len(names)
        reps = 6
        with torch.no_grad():
            for f in range(reps + 2):
                bd = batch_dict(f, dev_pts[f % NUM_CLOUDS])
                flush.zero_()
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
                evs[0].record()
                for k, m in enumerate(model.module_list):
                    bd = m(bd)
                    evs[k + 1].record()
                torch.cuda.synchronize()
                if f >= 2:
                    for k in range(len(names)):
                        acc[k] += evs[k].elapsed_time(evs[k + 1]) / reps
        print('stage times (eager, us/frame): ' + ', '.join('%s %.0f' % (n, 1000 * t) for n, t in zip(names, acc)) +
              ' | total %.0f' % (1000 * sum(acc)), file=sys.stderr)

    if args.modules and rank == 0:
        full = list(model.module_list)
        for nmod in range(1, len(full) + 1):
            model.module_list = full[:nmod]
            sp = dev_pts[0].clone()
            gg, _ = model.capture_graph(batch_dict(0, sp), warmup=1)
            ts = []
            for i in range(13):
                flush.zero_()
                sp.copy_(dev_pts[i % NUM_CLOUDS])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gg.replay(); e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ts.sort()
            print('graph replay of the first %d modules (.. %s): median %.0f us' % (nmod, type(full[nmod - 1]).__name__, 1000 * ts[len(ts) // 2]), file=sys.stderr)
        model.module_list = full
        return

    if not args.no_graph:
        static_pts = dev_pts[0].clone()
        ops.reset_launch_count()
        g, out = model.capture_graph(batch_dict(0, static_pts), warmup=0)
        graph_state.update(g=g, out=out, launches_per_step=ops.launch_count())
        graph_state['in'] = static_pts
        for i in range(NUM_CLOUDS):                         # replay sanity: same detections as the eager path
            step_resident(i)
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    total_ms, launches = timed(step_resident, args.steps, args.warmup)
    sampler.stop_flag = True
    e2e_ms, _ = timed(step_e2e, args.steps, args.warmup)
    with torch.no_grad():                                   # overflow check of the resident-arm configuration
        model.post_processing(step_resident(0))
    frames = args.steps * args.batch * world
    value = frames / (total_ms / 1000.0)
    e2e = frames / (e2e_ms / 1000.0)

    # ---- roofline of the dominant kernel family: the sparse convolution layers, timed live with CUDA events
    def step_eager(i):
        with torch.no_grad():
            return model.forward_device(batch_dict(i, dev_pts[i % NUM_CLOUDS]))
    roof = sparse_conv_roofline(model, step_eager, args, dev)

    line = {
        'metric': 'Waymo-shape frames/sec', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': ('fp32' if args.mode == 'fp32' and args.sp_mode == 'fp32' else
                                      'fp32 storage and accumulation; products: sparse convs %s, dense BEV/head convs %s' % (args.sp_mode, args.mode)),
        'data': 'synthetic',
        'config': {'workload': 'CenterPoint 1-sweep %s, synthetic 180K-pt Waymo-range cloud, fp32, batch %d/GPU'
                               % (args.backbone, args.batch),
                   'sparse_conv_mode': args.sp_mode, 'dense_conv_mode': args.mode, 'parallelism': 'frames sharded dp%d' % world,
                   'l2': 'flushed (256 MiB write) between steps, outside the timed intervals',
                   'launch': 'eager' if args.no_graph else 'CUDA graph replay of CenterPoint.forward_device',
                   'detections_last_step': int(e2e_state['boxes'])},
        'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': int(host_pts[0].numel() * 4),
                'd2h_bytes_per_step': int(e2e_state['d2h'])},          # boxes of every frame + the count read, last step
        'gpu_launches': launches,
        'clocks': sampler.summary(),
        'roofline': roof,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sparse_conv_roofline(model, step_fn, args, dev):
    """achieved = algorithmic bytes of all sparse-conv launches of one frame / their summed duration (CUDA events on
    the launching stream); algorithmic bytes per layer = (N_in*Cin + N_out*Cout)*4 + pairs*8 + K*Cin*Cout*4
    (+ N_out*Cout*4 with a residual) -- SURVEY.md §8d."""
    import torch
    from detzero_b200 import ops
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, which = json.load(open(peaks_path))['hbm_gbs'], 'measured'
    else:
        peak, which = 6650.0, 'fallback'
    rec = ops.enable_spconv_trace(True)
    step_fn(0)
    torch.cuda.synchronize()
    ops.enable_spconv_trace(False)
    tot_bytes, tot_ms, tot_flops = 0.0, 0.0, 0.0
    for r in rec:
        if r['nbr'].shape[1] == 32 and r['nbr'].shape[0] != r['K']:      # tensor-core launch: row-major table (cap, 32)
            valid = r['nbr'][:r['n_out'], :r['K']].t() >= 0
            if r['row_order'] is not None:
                valid = valid[:, r['row_order'][:r['n_out']].long()]        # tile order (for the offsets/tile statistic)
        else:
            valid = r['nbr'][:, :r['n_out']] >= 0
        pairs = int(valid.sum().item())
        if args.layer_times:
            T = (r['n_out'] + 127) // 128
            v = torch.zeros((valid.shape[0], T * 128), dtype=torch.bool, device=valid.device)
            v[:, :r['n_out']] = valid
            touched = int(v.view(valid.shape[0], T, 128).any(2).sum().item())
            print('spconv layer K=%d cin=%d cout=%d n_out=%d pairs=%d offsets/tile=%.1f scheduled=%s  %.1f us' % (
                r['K'], r['cin'], r['cout'], r['n_out'], pairs, touched / max(T, 1), r['row_order'] is not None,
                1000 * r['start'].elapsed_time(r['end'])), file=sys.stderr)
        K, cin, cout = r['K'], r['cin'], r['cout']
        b = (r['n_in'] * cin + r['n_out'] * cout) * 4 + pairs * 8 + K * cin * cout * 4 + (r['n_out'] * cout * 4 if r['residual'] else 0)
        tot_bytes += b
        tot_flops += 2.0 * pairs * cin * cout
        tot_ms += r['start'].elapsed_time(r['end'])
    achieved = tot_bytes / (tot_ms / 1000.0) / 1e9 if tot_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if os.path.exists(tpath) and args.sp_mode == 'tf32' and args.backbone == 'VoxelBackBone8x':
        tj = json.load(open(tpath))              # from the committed ncu --set full captures (same launches, cold caches)
        if args.batch == tj.get('frames_per_step'):
            traffic = tj['sparse_conv_dram_bytes_per_step']
        elif args.batch == 1:
            traffic = tj['batch1']['sparse_conv_dram_bytes_per_frame']
    return {'bound': 'hbm', 'kernel': 'k_spconv (all %d sparse-conv launches of one step)' % len(rec), 'achieved': achieved,
            'peak': peak, 'peak_source': which, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
            'algorithmic_bytes_per_step': tot_bytes, 'algorithmic_flops_per_step': tot_flops, 'ms_per_step': tot_ms,
            'frames_per_step': args.batch}


def cpu_threads():
    """threads the CPU arm uses: the oracle port (numpy rulebooks + torch gather/mm/index_add + oneDNN convs) stops scaling
    -- and on a 128-core host gets ~30x SLOWER from oversubscription -- beyond ~16 threads"""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get('DZ_CPU_THREADS', '16'))))


def cpu_baseline(args):
    import torch
    cores = cpu_threads()
    torch.set_num_threads(cores)
    run = cpu_frame_fn(args.backbone)
    run(0)
    t0 = time.perf_counter()
    n = 0
    while n < 2 or (time.perf_counter() - t0 < 12.0 and n < 8):
        run(n)
        n += 1
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': '%d frames of the same workload (oracle port: C voxelizer + spconv-Native restatement + torch-CPU convs)' % n}


if __name__ == '__main__':
    main()
