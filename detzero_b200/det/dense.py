"""HeightCompression, BaseBEVBackbone and CenterHead with the reference's names, ctor kwargs, module trees
(=> identical state-dict keys) and batch_dict contract:
  detection/detzero_det/models/centerpoint_modules/height_compression.py:4-25
  detection/detzero_det/models/centerpoint_modules/backbone2d.py:6-120
  detection/detzero_det/models/centerpoint_modules/center_head.py:14-48,50-100,315-368,440-488
running on the NHWC conv2d / decode / NMS kernels of libdetzero_b200.

Dense maps keep their logical NCHW shape in ``batch_dict`` but live in channels_last (NHWC) memory, so the
reference-visible shapes are unchanged while the kernels see contiguous channel vectors."""
import copy

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops
from ..spconv.pytorch import fold_bn


def _nhwc(x):
    """logical NCHW tensor -> contiguous (B,H,W,C) view (converts only if the memory is not channels_last)"""
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


def _nchw_view(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


def _cached(key_obj, tensors, fn):
    """kernel-layout copies are cached ON THE MODULE (a global dict keyed by id() can hand a freed model's weights to a new
    model that recycles the id and the allocator address)"""
    mod, sub = key_obj if isinstance(key_obj, tuple) else (key_obj, None)
    ver = tuple((t._version, t.data_ptr(), t.device) for t in tensors if t is not None)
    cache = mod.__dict__.setdefault('_dz_pack', {})
    hit = cache.get(sub)
    if hit is not None and hit[0] == ver:
        return hit[1]
    with torch.no_grad():
        val = fn()
    cache[sub] = (ver, val)
    return val


def pack_conv_weight(conv, mode):
    return _cached((conv, mode), [conv.weight], lambda: ops.pack_conv_weight(conv.weight, mode))


def pack_deconv_weight(deconv, mode):
    return _cached((deconv, mode), [deconv.weight], lambda: ops.pack_deconv_weight(deconv.weight, mode))


def run_conv_stack(seq, x, mode, out=None, out_coff=0):
    """Execute an nn.Sequential made of [ZeroPad2d] Conv2d|ConvTranspose2d [BatchNorm2d] [ReLU] groups on NHWC
    input with fused epilogues.  The last group may write into a channel slice of ``out`` (fused concat)."""
    mods = list(seq.children())
    i, pad_extra = 0, 0
    groups = []
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.ZeroPad2d):
            pad_extra = int(m.padding[0])
            i += 1
            continue
        assert isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)), type(m)
        bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
        j = i + (2 if bn is not None else 1)
        relu = j < len(mods) and isinstance(mods[j], nn.ReLU)
        groups.append((m, bn, relu, pad_extra))
        pad_extra = 0
        i = j + (1 if relu else 0)
    for gi, (m, bn, relu, pad_extra) in enumerate(groups):
        last = gi == len(groups) - 1
        if bn is not None:
            scale, shift = fold_bn(bn, m.bias)
        else:
            scale, shift = None, (None if m.bias is None else m.bias.detach().float())
        o, off = (out, out_coff) if last else (None, 0)
        if isinstance(m, nn.ConvTranspose2d):
            assert m.kernel_size[0] == m.stride[0] and m.padding[0] == 0
            x = ops.deconv2d(x, pack_deconv_weight(m, mode), (m.kernel_size[0], m.in_channels, m.out_channels), scale, shift,
                             relu, out=o, out_coff=off, mode=mode)
        else:
            x = ops.conv2d(x, pack_conv_weight(m, mode), (m.kernel_size[0], m.kernel_size[1], m.in_channels, m.out_channels),
                           int(m.stride[0]), int(m.padding[0]) + pad_extra, scale, shift, relu, out=o, out_coff=off, mode=mode)
    return x


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES

    def forward(self, batch_dict):
        t = batch_dict['encoded_spconv_tensor']
        D, H, W = t.spatial_shape
        nhwc = ops.sparse_to_bev(t._feat, t._idx, t._count, t._cap, t.batch_size, D, H, W, planes=t._planes)   # channel = c*D + z
        batch_dict['spatial_features'] = _nchw_view(nhwc)          # (N, C*D, H, W), channels_last memory
        batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
        return batch_dict


class BaseBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        self.mode = _lib.MODES[model_cfg.get('COMPUTE_MODE', 'fp32')]
        if self.model_cfg.get('LAYER_NUMS', None) is not None:
            layer_nums, layer_strides, num_filters = (list(self.model_cfg.LAYER_NUMS), list(self.model_cfg.LAYER_STRIDES),
                                                      list(self.model_cfg.NUM_FILTERS))
            assert len(layer_nums) == len(layer_strides) == len(num_filters)
        else:
            layer_nums = layer_strides = num_filters = []
        if self.model_cfg.get('UPSAMPLE_STRIDES', None) is not None:
            num_upsample_filters, upsample_strides = list(self.model_cfg.NUM_UPSAMPLE_FILTERS), list(self.model_cfg.UPSAMPLE_STRIDES)
            assert len(upsample_strides) == len(num_upsample_filters)
        else:
            upsample_strides = num_upsample_filters = []
        num_levels = len(layer_nums)
        c_in_list = [input_channels, *num_filters[:-1]]
        bn = lambda c: nn.BatchNorm2d(c, eps=1e-3, momentum=0.01)
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(num_levels):
            layers = [nn.ZeroPad2d(1),
                      nn.Conv2d(c_in_list[idx], num_filters[idx], kernel_size=3, stride=layer_strides[idx], padding=0, bias=False),
                      bn(num_filters[idx]), nn.ReLU()]
            for _ in range(layer_nums[idx]):
                layers += [nn.Conv2d(num_filters[idx], num_filters[idx], kernel_size=3, padding=1, bias=False),
                           bn(num_filters[idx]), nn.ReLU()]
            self.blocks.append(nn.Sequential(*layers))
            if len(upsample_strides) > 0:
                stride = upsample_strides[idx]
                if stride >= 1:
                    up = nn.ConvTranspose2d(num_filters[idx], num_upsample_filters[idx], upsample_strides[idx],
                                            stride=upsample_strides[idx], bias=False)
                else:
                    k = int(np.round(1 / stride))
                    up = nn.Conv2d(num_filters[idx], num_upsample_filters[idx], k, stride=k, bias=False)
                self.deblocks.append(nn.Sequential(up, bn(num_upsample_filters[idx]), nn.ReLU()))
        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > num_levels:
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(c_in, c_in, upsample_strides[-1], stride=upsample_strides[-1], bias=False), bn(c_in), nn.ReLU()))
        self.num_bev_features = c_in
        self._up_channels = list(num_upsample_filters)

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError('training through the fused NHWC conv path is a next row (SURVEY.md §8f)')
        sf = data_dict['spatial_features']
        x = _nhwc(sf)
        B, H0 = x.shape[0], x.shape[1]
        n_cat = min(len(self.blocks), len(self.deblocks)) if len(self.deblocks) > 0 else 0
        cat = None
        ups = []
        for i in range(len(self.blocks)):
            x = run_conv_stack(self.blocks[i], x, self.mode)
            stride = int(H0 / x.shape[1])
            data_dict['spatial_features_%dx' % stride] = _nchw_view(x)
            if len(self.deblocks) > 0:
                if cat is None:
                    # all deblock outputs share the spatial size of the first one; write them straight into the
                    # concatenated (B,H,W,sum C) tensor (fuses torch.cat, backbone2d.py:107-108)
                    up0 = self.deblocks[i][0]
                    s0 = up0.stride[0] if isinstance(up0, nn.ConvTranspose2d) else 1.0 / up0.stride[0]
                    Hc, Wc = int(round(x.shape[1] * s0)), int(round(x.shape[2] * s0))
                    cat = torch.empty((B, Hc, Wc, sum(self._up_channels[:n_cat])), dtype=torch.float32, device=x.device)
                off = sum(self._up_channels[:i])
                run_conv_stack(self.deblocks[i], x, self.mode, out=cat, out_coff=off)
            else:
                ups.append(x)
        y = cat if cat is not None else (ups[0] if len(ups) == 1 else torch.cat(ups, dim=3))
        if len(self.deblocks) > len(self.blocks):
            y = run_conv_stack(self.deblocks[-1], y, self.mode)
        data_dict['spatial_features_2d'] = _nchw_view(y)
        return data_dict


class SeparateHead(nn.Module):
    """center_head.py:14-48 (module tree and initialisation)"""

    def __init__(self, input_channels, sep_head_dict, init_bias=-2.19, use_bias=False):
        super().__init__()
        self.sep_head_dict = sep_head_dict
        for cur_name in self.sep_head_dict:
            output_channels = self.sep_head_dict[cur_name]['out_channels']
            num_conv = self.sep_head_dict[cur_name]['num_conv']
            fc_list = []
            for _ in range(num_conv - 1):
                fc_list.append(nn.Sequential(
                    nn.Conv2d(input_channels, input_channels, kernel_size=3, stride=1, padding=1, bias=use_bias),
                    nn.BatchNorm2d(input_channels), nn.ReLU()))
            fc_list.append(nn.Conv2d(input_channels, output_channels, kernel_size=3, stride=1, padding=1, bias=True))
            fc = nn.Sequential(*fc_list)
            if 'hm' in cur_name:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight.data)
                        if m.bias is not None:
                            nn.init.constant_(m.bias, 0)
            self.__setattr__(cur_name, fc)


class CenterHead(nn.Module):
    def __init__(self, model_cfg, input_channels, num_class, class_names, grid_size, point_cloud_range, voxel_size,
                 tta=False, predict_boxes_when_training=True):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.grid_size = grid_size
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.voxel_size = [float(v) for v in voxel_size]
        self.TTA = tta
        self.mode = _lib.MODES[model_cfg.get('COMPUTE_MODE', 'fp32')]
        self.iou_weight = self.model_cfg.get('IOU_WEIGHT', 0)
        self.feature_map_stride = self.model_cfg.TARGET_ASSIGNER_CONFIG.get('FEATURE_MAP_STRIDE', None)
        self.class_names = class_names
        self.class_names_each_head = []
        self.class_id_mapping_each_head = []
        for cur_class_names in self.model_cfg.CLASS_NAMES_EACH_HEAD:
            self.class_names_each_head.append([x for x in cur_class_names if x in class_names])
            self.class_id_mapping_each_head.append([self.class_names.index(x) for x in cur_class_names if x in class_names])
        total_classes = sum(len(x) for x in self.class_names_each_head)
        assert total_classes == len(self.class_names), f'class_names_each_head={self.class_names_each_head}'
        sc = self.model_cfg.SHARED_CONV_CHANNEL
        use_bias = self.model_cfg.get('USE_BIAS_BEFORE_NORM', False)
        self.shared_conv = nn.Sequential(nn.Conv2d(input_channels, sc, 3, stride=1, padding=1, bias=use_bias),
                                         nn.BatchNorm2d(sc), nn.ReLU())
        self.heads_list = nn.ModuleList()
        self.separate_head_cfg = self.model_cfg.SEPARATE_HEAD_CFG
        self.head_order = list(self.separate_head_cfg.HEAD_ORDER) + ['hm']
        for cur_class_names in self.class_names_each_head:
            cur_head_dict = copy.deepcopy(dict(self.separate_head_cfg.HEAD_DICT))
            cur_head_dict['hm'] = dict(out_channels=len(cur_class_names), num_conv=self.model_cfg.NUM_HM_CONV)
            self.heads_list.append(SeparateHead(sc, cur_head_dict, init_bias=-2.19, use_bias=use_bias))
        self.predict_boxes_when_training = predict_boxes_when_training
        self.forward_ret_dict = {}

    # ---- fused branch weights -----------------------------------------------------------------------------
    def _fused_head(self, head):
        """All branches of one SeparateHead as two convs: stage 1 = concatenated 64->64 convs (+BN+ReLU) as one
        64->(64*nb) conv; stage 2 = the per-branch 64->out convs as one block-diagonal (64*nb)->sum(out) conv."""
        names = self.head_order
        params = []
        for n in names:
            fc = getattr(head, n)
            assert len(fc) == 2, 'fused path expects num_conv == 2 (shipped configs)'
            params += [fc[0][0].weight, fc[0][0].bias, fc[0][1].weight, fc[0][1].bias, fc[0][1].running_mean,
                       fc[0][1].running_var, fc[1].weight, fc[1].bias]

        mode = self.mode

        def build():
            sc = getattr(head, names[0])[0][0].in_channels
            w1, s1, b1 = [], [], []
            outs = [getattr(head, n)[1].out_channels for n in names]
            tot = sum(outs)
            tot_pad = (tot + 3) // 4 * 4
            w2 = torch.zeros((tot_pad, sc * len(names), 3, 3), dtype=torch.float32, device=params[0].device)   # OIHW
            b2 = torch.zeros((tot_pad,), dtype=torch.float32, device=params[0].device)
            off, layout = 0, {}
            for bi, n in enumerate(names):
                fc = getattr(head, n)
                w1.append(fc[0][0].weight.detach().float())                                 # (sc, sc, 3, 3)
                sc_, sh_ = fold_bn(fc[0][1], fc[0][0].bias)
                s1.append(sc_); b1.append(sh_)
                w2[off:off + outs[bi], bi * sc:(bi + 1) * sc] = fc[1].weight.detach().float()   # block diagonal
                b2[off:off + outs[bi]] = fc[1].bias.detach().float()
                layout[n] = off
                off += outs[bi]
            w1 = torch.cat(w1, dim=0)                                                      # (sc*nb, sc, 3, 3)
            return (ops.pack_conv_weight(w1, mode), (3, 3, sc, sc * len(names)), torch.cat(s1).contiguous(),
                    torch.cat(b1).contiguous(), ops.pack_conv_weight(w2, mode), (3, 3, sc * len(names), tot_pad), b2, layout)
        return _cached((head, self.mode), params, build)

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError('CenterHead training (assign_targets/losses) is a next row (SURVEY.md §8f)')
        x = _nhwc(data_dict['spatial_features_2d'])
        x = run_conv_stack(self.shared_conv, x, self.mode)
        B, H, W, _ = x.shape
        pred_dicts, padded, counts = [], [], []
        post = self.model_cfg.POST_PROCESSING
        for idx, head in enumerate(self.heads_list):
            w1, k1, s1, b1, w2, k2, b2, layout = self._fused_head(head)
            h1 = ops.conv2d(x, w1, k1, 1, 1, s1, b1, True, mode=self.mode)
            hm = ops.conv2d(h1, w2, k2, 1, 1, None, b2, False, mode=self.mode)           # (B,H,W,tot_pad)
            nchw = _nchw_view(hm)
            outs = {n: nchw[:, layout[n]:layout[n] + getattr(head, n)[1].out_channels] for n in self.head_order}
            pred_dicts.append(outs)
            # NOTE: the reference also runs assign_targets here at test time (`if self.training or True`,
            # center_head.py:448-453) -- pure host-side overhead that does not influence predictions; skipped.
            K = int(post.MAX_OBJ_PER_SAMPLE)
            boxes, scores, labels, d_n = ops.centerhead_decode(
                hm, {k: layout.get(k, 0) for k in ('center', 'center_z', 'dim', 'rot', 'iou', 'hm')},
                len(self.class_names_each_head[idx]), K, self.point_cloud_range, self.voxel_size,
                self.feature_map_stride, list(post.POST_CENTER_LIMIT_RANGE), float(post.SCORE_THRESH),
                self.iou_weight > 0 and 'iou' in layout)
            assert post.NMS_CONFIG.NMS_TYPE == 'nms_gpu', 'only the shipped rotated NMS is built'
            mapping = self.class_id_mapping_each_head[idx]
            assert mapping == list(range(mapping[0], mapping[0] + len(mapping))), 'non-contiguous class mapping'
            # 'gather_slab' = (boxes (B,500,9), counts (B,)) views into the send buffer of the per-sequence box gather
            # (dist.SequenceGather.slot): the NMS writes its result straight where the collective reads it
            slab = data_dict.get('gather_slab') if len(self.heads_list) == 1 else None
            out, d_out = ops.nms_bev(boxes, scores, labels, d_n, float(post.NMS_CONFIG.NMS_THRESH),
                                     int(post.NMS_CONFIG.NMS_POST_MAXSIZE), label_offset=mapping[0] + 1,
                                     out=None if slab is None else slab[0], d_out_n=None if slab is None else slab[1])
            padded.append(out)
            counts.append(d_out)
        self.forward_ret_dict['pred_dicts'] = pred_dicts
        # fixed-shape device result: rows [x,y,z,dx,dy,dz,heading,score,label(1-based)]
        data_dict['final_boxes_padded'] = padded[0] if len(padded) == 1 else torch.cat(padded, dim=1)
        data_dict['final_boxes_count'] = counts[0] if len(counts) == 1 else torch.stack(counts, dim=1)
        return data_dict

    @staticmethod
    def boxes_to_dicts(padded, counts):
        """host materialisation of final_box_dicts (center_head.py:363-367): ONE device->host read of the counts"""
        cnt = counts.tolist()
        ret = []
        for b, n in enumerate(cnt):
            if isinstance(n, list):                                # multi-head: concatenate per-head prefixes
                per = padded.shape[1] // len(n)
                rows = torch.cat([padded[b, h * per:h * per + nh] for h, nh in enumerate(n)], dim=0)
            else:
                rows = padded[b, :n]
            ret.append({'pred_boxes': rows[:, :7], 'pred_scores': rows[:, 7], 'pred_labels': rows[:, 8].long()})
        return ret
