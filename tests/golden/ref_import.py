"""Import the UNMODIFIED reference modules from /root/reference in the build container (CPU), with stand-ins for
what is absent here: spconv (-> oracle.spconv_ref shim), the compiled CUDA ops (-> oracle C code), easydict, and
``.cuda()`` (identity on this GPU-less box).  Used only by make_golden.py; never on the GPU box."""
import os
import sys
import types

import torch

REF = '/root/reference'


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def install():
    assert os.path.isdir(REF), 'reference not mounted'
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import oracle
    from oracle import spconv_ref
    spconv_ref.install_shim()
    # packages whose __init__ needs generated version.py / drags in unrelated ops: expose the dirs only
    _pkg('detzero_utils', REF + '/utils/detzero_utils')
    _pkg('detzero_utils.ops', REF + '/utils/detzero_utils/ops')
    _pkg('detzero_utils.ops.iou3d_nms', REF + '/utils/detzero_utils/ops/iou3d_nms')
    _pkg('detzero_utils.ops.roiaware_pool3d', REF + '/utils/detzero_utils/ops/roiaware_pool3d')
    _pkg('detzero_det', REF + '/detection/detzero_det')
    _pkg('detzero_det.models', REF + '/detection/detzero_det/models')
    _pkg('detzero_det.models.centerpoint_modules', REF + '/detection/detzero_det/models/centerpoint_modules')
    _pkg('detzero_det.utils', REF + '/detection/detzero_det/utils')
    _pkg('detzero_refine', REF + '/refining/detzero_refine')
    _pkg('detzero_refine.models', REF + '/refining/detzero_refine/models')
    _pkg('detzero_refine.utils', REF + '/refining/detzero_refine/utils')
    # compiled ops -> oracle
    nms = types.ModuleType('detzero_utils.ops.iou3d_nms.iou3d_nms_cuda')

    def nms_gpu(boxes, keep, thresh):
        k = oracle.nms_bev_sorted(boxes.detach().cpu().numpy(), thresh)
        keep[:len(k)] = torch.from_numpy(k)
        return len(k)
    nms.nms_gpu = nms_gpu
    sys.modules['detzero_utils.ops.iou3d_nms.iou3d_nms_cuda'] = nms
    sys.modules['detzero_utils.ops.roiaware_pool3d.roiaware_pool3d_cuda'] = types.ModuleType('roiaware_pool3d_cuda')
    # easydict stand-in
    if 'easydict' not in sys.modules:
        from detzero_b200.config import AttrDict
        ed = types.ModuleType('easydict')
        ed.EasyDict = AttrDict
        sys.modules['easydict'] = ed
    for missing in ('SharedArray', 'tensorboardX'):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                sys.modules[missing] = types.ModuleType(missing)
    # no GPU here: .cuda() is the identity
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
