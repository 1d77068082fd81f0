"""Golden vectors for the refiner: the reference's own PositionTransformer / GeometryTransformer / ConfidencePointnet
(refining/detzero_refine/models/modules/*.py, unmodified) run on CPU with seeded weights and inputs."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_import  # noqa: E402

ref_import.install()
from oracle import refine_inputs as ri  # noqa: E402
from oracle import weights  # noqa: E402

SEED = 4321


def golden_refine():
    from detzero_refine.models.modules.confidence_pointnet import ConfidencePointnet
    from detzero_refine.models.modules.geometry_transformer import GeometryTransformer
    from detzero_refine.models.modules.position_transformer import PositionTransformer
    out, keys = {}, {}
    torch.manual_seed(0)
    prm = PositionTransformer(ri.prm_cfg(), 32, 32).eval()
    weights.load_seeded(prm, SEED)
    keys['PositionTransformer'] = [(k, list(v.shape)) for k, v in prm.state_dict().items()]
    with torch.no_grad():
        d = prm(ri.prm_inputs(SEED))
    out['prm.query'] = d['query'].numpy()                       # (B,256,200) encoder output
    out['prm.memory_sum'] = np.array([d['memory'].double().sum().item(), d['memory'].double().abs().sum().item()])
    for k in ('center_reg', 'heading_cls', 'heading_reg'):
        out['prm.' + k] = prm.preds_dict[k].numpy()
    out['prm.batch_box_preds'] = d['batch_box_preds'].numpy()
    grm = GeometryTransformer(ri.grm_cfg(), 11, 4).eval()
    weights.load_seeded(grm, SEED + 1)
    keys['GeometryTransformer'] = [(k, list(v.shape)) for k, v in grm.state_dict().items()]
    with torch.no_grad():
        d = grm(ri.grm_inputs(SEED + 1))
    out['grm.geometry_cls'] = grm.preds_dict['geometry_cls'].numpy()
    out['grm.geometry_reg'] = grm.preds_dict['geometry_reg'].numpy()
    out['grm.batch_box_preds'] = d['batch_box_preds'].numpy()
    crm = ConfidencePointnet(ri.crm_cfg(), 32, 32).eval()
    weights.load_seeded(crm, SEED + 2)
    keys['ConfidencePointnet'] = [(k, list(v.shape)) for k, v in crm.state_dict().items()]
    with torch.no_grad():
        d = crm(ri.crm_inputs(SEED + 2))
    out['crm.pred_score'] = d['pred_score'].numpy()
    np.savez_compressed(os.path.join(HERE, 'refine.npz'), **out)
    with open(os.path.join(HERE, 'refine_state_dict_keys.json'), 'w') as f:
        json.dump(keys, f, indent=0)
    print('refine.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    golden_refine()
