"""Call-compatible mirror of mmdet/models/roi_extractors/single_level.py with the level loop fused away.

The reference maps every RoI to an FPN level (`map_roi_levels`, single_level.py:54-73) and then loops over the
levels: boolean mask -> gather rois -> one RoIAlign launch -> masked scatter into `roi_feats`
(single_level.py:89-107) -- 4 kernel launches and 8 index ops with host synchronisation (`inds.any()`) per
call, twice per image in the OBB detectors (bbox and rbbox extractors, mmdet/models/detectors/test_mixins.py:
269-322,352-380).  Here the per-RoI level id is an argument of the kernel (aidet_rroi_align_fwd_f32 takes all
level pointers), so one launch serves P2-P5 and no gather / scatter exists; the backward is one gather pass
over all levels too.  Rotated RoIs (k,6) = [batch, cx, cy, w, h, theta] are accepted with
`roi_layer=dict(type='RoIAlignRotated', ...)`.
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import ops
from ...ops import functional as F
from ...ops.roi_align.roi_align import RoIAlign, RoIAlignRotated, _nhwc, _variant


class _MultiLevelRoIAlign(Function):
    """All FPN levels in one forward launch / one gather-backward pass.  Gradient to the feature maps only
    (roi_align.py:60-73)."""

    @staticmethod
    def forward(ctx, rois, lvls, out_size, scales, sample_num, variant, *feats):
        if not all(f.is_cuda for f in feats):
            raise NotImplementedError          # roi_align.py:41-42: no CPU path
        ctx.scales, ctx.sample_num, ctx.variant = scales, sample_num, variant
        ctx.sizes = [f.size() for f in feats]
        ctx.dtypes = [f.dtype for f in feats]
        ctx.save_for_backward(rois, lvls)
        nhwc = [_nhwc(f.float()) for f in feats]
        out = F.rroi_align_forward(nhwc, rois, scales, out_size, sample_num, variant, roi_level=lvls)
        return out.permute(0, 3, 1, 2).to(feats[0].dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, lvls = ctx.saved_tensors
        go = grad_output.float().permute(0, 2, 3, 1).contiguous()
        dev = grad_output.device
        c = ctx.sizes[0][1]
        if ctx.sample_num > 0 and c % 4 == 0:
            gfs = [torch.empty((n, h, w, ch), dtype=torch.float32, device=dev) for (n, ch, h, w) in ctx.sizes]
            F.rroi_align_backward_gather(go, gfs, rois, ctx.scales, ctx.sample_num, ctx.variant, roi_level=lvls)
        else:
            gfs = [torch.zeros((n, h, w, ch), dtype=torch.float32, device=dev) for (n, ch, h, w) in ctx.sizes]
            F.rroi_align_backward(go, gfs, rois, ctx.scales, ctx.sample_num, ctx.variant, roi_level=lvls)
        grads = tuple(g.permute(0, 3, 1, 2).to(dt) if need else None
                      for g, dt, need in zip(gfs, ctx.dtypes, ctx.needs_input_grad[6:]))
        return (None, None, None, None, None, None) + grads


class SingleRoIExtractor(nn.Module):
    """Extract RoI features from a single level feature map (single_level.py:11-107).

    Args:
        roi_layer (dict): RoI layer type and arguments, e.g. dict(type='RoIAlign', out_size=7, sample_num=2)
            or dict(type='RoIAlignRotated', ...); looked up by name in `ops` like the reference (:47-49).
        out_channels (int): Output channels of RoI layers.
        featmap_strides (list[int]): Strides of input feature maps.
        finest_scale (int): Scale threshold of mapping to level 0.
    """

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super(SingleRoIExtractor, self).__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        """int: Input feature map levels."""
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert hasattr(ops, layer_type)
        layer_cls = getattr(ops, layer_type)
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def map_roi_levels(self, rois, num_levels):
        """Level index (0-based) of each RoI by scale (single_level.py:54-73):
        scale < finest_scale * 2 -> 0, < * 4 -> 1, < * 8 -> 2, else 3.  (k,5) rois use the reference's `+1` box
        sides; rotated (k,6) rois use their own (w + 1, h + 1), i.e. the same box model."""
        if rois.size(1) == 5:
            scale = torch.sqrt((rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1))
        else:
            scale = torch.sqrt((rois[:, 3] + 1) * (rois[:, 4] + 1))
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def roi_rescale(self, rois, scale_factor):
        if rois.size(1) == 6:
            new = rois.clone()
            new[:, 3] = (rois[:, 3] + 1) * scale_factor - 1
            new[:, 4] = (rois[:, 4] + 1) * scale_factor - 1
            return new
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        w = rois[:, 3] - rois[:, 1] + 1
        h = rois[:, 4] - rois[:, 2] + 1
        new_w = w * scale_factor
        new_h = h * scale_factor
        x1 = cx - new_w * 0.5 + 0.5
        x2 = cx + new_w * 0.5 - 0.5
        y1 = cy - new_h * 0.5 + 0.5
        y2 = cy + new_h * 0.5 - 0.5
        return torch.stack((rois[:, 0], x1, y1, x2, y2), dim=-1)

    def _fusable(self):
        l0 = self.roi_layers[0]
        if not isinstance(l0, (RoIAlign, RoIAlignRotated)) or getattr(l0, 'use_torchvision', False):
            return False
        return all(type(l) is type(l0) and l.out_size == l0.out_size and l.sample_num == l0.sample_num
                   and l.aligned == l0.aligned and not getattr(l, 'use_torchvision', False) for l in self.roi_layers)

    def forward(self, feats, rois, roi_scale_factor=None):
        if len(feats) == 1:
            return self.roi_layers[0](feats[0], rois)
        num_levels = len(feats)
        target_lvls = self.map_roi_levels(rois, num_levels)
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        if self._fusable() and feats[0].is_cuda:
            l0 = self.roi_layers[0]
            assert rois.dim() == 2 and rois.size(1) == (6 if isinstance(l0, RoIAlignRotated) else 5)
            scales = [l.spatial_scale for l in self.roi_layers][:num_levels]
            return _MultiLevelRoIAlign.apply(rois, target_lvls.int(), l0.out_size, scales, l0.sample_num,
                                             _variant(l0.aligned, legacy=not l0.aligned), *feats[:num_levels])
        # reference loop (single_level.py:96-107), kept for layers the fused kernel does not serve
        out_size = self.roi_layers[0].out_size
        roi_feats = feats[0].new_zeros(rois.size(0), self.out_channels, *out_size)
        for i in range(num_levels):
            inds = target_lvls == i
            if inds.any():
                roi_feats[inds] = self.roi_layers[i](feats[i], rois[inds, :])
        return roi_feats
