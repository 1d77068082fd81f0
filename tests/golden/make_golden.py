"""Generate golden vectors by running the reference's OWN Python modules (unmodified, imported from /root/reference)
on CPU in the build container.  The reference cannot travel to the GPU box, so the outputs are committed as small
fixtures (tests/golden/*.npz) together with this script.  Weights and inputs are not stored: both sides regenerate them
from oracle/weights.py (a seeded function of key name + shape).

    python tests/golden/make_golden.py            # needs /root/reference

What runs from the reference: backbone3d.py (VoxelBackBone8x, VoxelResBackBone8x, SparseBasicBlock, post_act_block) on
the oracle's spconv shim; vfe.py MeanVFE; height_compression.py; backbone2d.py; center_head.py (forward incl.
generate_predicted_boxes -> centernet_utils.decode_bbox_from_heatmap -> model_nms_utils.class_agnostic_nms, with the
compiled NMS op replaced by the oracle's C restatement); refiner modules (see make_golden_refine below).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402

ref_import.install()

import oracle  # noqa: E402
from oracle import weights  # noqa: E402
from tests import util  # noqa: E402

TINY_RANGE = [-4.8, -4.8, -2, 4.8, 4.8, 4]
GRID = np.array([96, 96, 40])
SEED = 1234


def tiny_batch():
    clouds = [util.clustered_cloud(6000, 101, TINY_RANGE), util.clustered_cloud(3500, 102, TINY_RANGE)]
    vox = oracle.Point2VoxelCPU3d(util.VOXEL, TINY_RANGE, 5, 5, 200000)
    v, c, n = [], [], []
    for b, p in enumerate(clouds):
        vv, cc, nn = vox.point_to_voxel(p)
        v.append(vv); n.append(nn); c.append(np.pad(cc, ((0, 0), (1, 0)), constant_values=b))
    return np.concatenate(v), np.concatenate(c).astype(np.int32), np.concatenate(n)


def golden_detector():
    from detzero_det.models.centerpoint_modules import backbone2d, backbone3d, center_head, height_compression, vfe
    out = {}
    voxels, coords, num = tiny_batch()
    bd0 = {'voxels': torch.from_numpy(voxels), 'voxel_num_points': torch.from_numpy(num).float(),
           'voxel_coords': torch.from_numpy(coords).float(), 'batch_size': 2}
    bd0 = vfe.MeanVFE(None, 5)(bd0)
    out['mean_vfe'] = bd0['voxel_features'].numpy()
    cfg = util.model_cfg()
    cfg.DENSE_HEAD.POST_PROCESSING.MAX_OBJ_PER_SAMPLE = 100     # reference _topk needs K <= H*W (12x12 map here)
    keys = {}
    for kind in ('VoxelBackBone8x', 'VoxelResBackBone8x'):
        m = getattr(backbone3d, kind)(cfg.BACKBONE_3D, 5, GRID).eval()
        weights.load_seeded(m, SEED)
        keys[kind] = [(k, list(v.shape)) for k, v in m.state_dict().items()]
        with torch.no_grad():
            bd = m(dict(bd0))
        t = bd['encoded_spconv_tensor']
        out[kind + '.out_idx'] = t.indices.numpy().astype(np.int32)
        out[kind + '.out_feat'] = t.features.numpy()
        for lv, st in bd['multi_scale_3d_features'].items():
            out['%s.%s.n' % (kind, lv)] = np.array([st.features.shape[0]])
            out['%s.%s.sum' % (kind, lv)] = np.array([st.features.double().sum().item(), st.features.double().abs().sum().item()])
        if kind == 'VoxelResBackBone8x':
            res_bd = bd
    # dense part on the Res backbone output
    hc = height_compression.HeightCompression(cfg.MAP_TO_BEV)
    bd = hc(res_bd)
    b2d = backbone2d.BaseBEVBackbone(cfg.BACKBONE_2D, 256).eval()
    weights.load_seeded(b2d, SEED + 1)
    keys['BaseBEVBackbone'] = [(k, list(v.shape)) for k, v in b2d.state_dict().items()]
    head = center_head.CenterHead(cfg.DENSE_HEAD, 512, 3, util.CLASS_NAMES, GRID, np.array(TINY_RANGE, np.float32), util.VOXEL,
                                  predict_boxes_when_training=False).eval()   # as CenterPoint builds it (centerpoint.py:106)
    weights.load_seeded(head, SEED + 2)
    keys['CenterHead'] = [(k, list(v.shape)) for k, v in head.state_dict().items()]
    with torch.no_grad():
        bd = b2d(bd)
        out['spatial_features_2d'] = bd['spatial_features_2d'].numpy()
        bd['gt_boxes'] = torch.zeros((2, 1, 8))
        bd = head(bd)
    for n, v in head.forward_ret_dict['pred_dicts'][0].items():
        out['head.' + n] = v.numpy()
    for b, d in enumerate(bd['final_box_dicts']):
        out['final.%d.boxes' % b] = d['pred_boxes'].numpy()
        out['final.%d.scores' % b] = d['pred_scores'].numpy()
        out['final.%d.labels' % b] = d['pred_labels'].numpy()
    np.savez_compressed(os.path.join(HERE, 'detector.npz'), **out)
    import json
    with open(os.path.join(HERE, 'state_dict_keys.json'), 'w') as f:
        json.dump(keys, f, indent=0)
    print('detector.npz:', {k: v.shape for k, v in out.items() if 'sum' not in k and '.n' not in k})


if __name__ == '__main__':
    golden_detector()
    try:
        from make_golden_refine import golden_refine
        golden_refine()
    except ImportError:
        pass
