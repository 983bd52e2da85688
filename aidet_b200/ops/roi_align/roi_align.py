"""Call-compatible mirror of mmdet/ops/roi_align/roi_align.py plus the rotated op.

`RoIAlign` / `roi_align` keep the reference surface (roi_align.py:9-152): ctor
`(out_size, spatial_scale, sample_num=0, use_torchvision=False, aligned=False)`, `.out_size`
pair, `forward(features NCHW, rois (K,5))`, gradient to features only, CPU tensors raise
`NotImplementedError` (roi_align.py:41-42).  `RoIAlignRotated` / `roi_align_rotated` take
rois `(K,6) = [batch, cx, cy, w, h, theta]` and reduce to the axis-aligned op at theta = 0.

Memory format: inputs are NCHW-logical.  The kernels run on channels-last storage; a
channels_last input is used in place, an NCHW-contiguous one is converted once.  Outputs
are (K,C,ph,pw) logical with channels-last strides.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import functional as F


def _variant(aligned, legacy):
    if legacy:
        return 0        # roi_align_kernel.cu (v1, +1)
    return 2 if aligned else 1      # roi_align_kernel_v2.cu


def _nhwc(features):
    """NCHW-logical tensor -> (N,H,W,C) contiguous view (no copy if already channels_last)."""
    return features.permute(0, 2, 3, 1).contiguous()


class _RoIAlignBase(Function):

    @staticmethod
    def _forward(ctx, features, rois, out_size, spatial_scale, sample_num, variant):
        out_h, out_w = _pair(out_size)
        assert isinstance(out_h, int) and isinstance(out_w, int)
        if not features.is_cuda:
            raise NotImplementedError
        ctx.spatial_scale = spatial_scale
        ctx.sample_num = sample_num
        ctx.variant = variant
        ctx.save_for_backward(rois)
        ctx.feature_size = features.size()
        feat = _nhwc(features.float())
        out = F.rroi_align_forward([feat], rois, [spatial_scale], (out_h, out_w), sample_num, variant)
        return out.permute(0, 3, 1, 2).to(features.dtype)

    @staticmethod
    def _backward(ctx, grad_output):
        rois = ctx.saved_tensors[0]
        assert ctx.feature_size is not None and grad_output.is_cuda
        n, c, h, w = ctx.feature_size
        grad_input = None
        if ctx.needs_input_grad[0]:
            go = grad_output.float().permute(0, 2, 3, 1).contiguous()
            if ctx.sample_num > 0 and c % 4 == 0:
                # deterministic gather backward: writes every pixel once, no zero-fill, no atomics
                gf = torch.empty((n, h, w, c), dtype=torch.float32, device=grad_output.device)
                F.rroi_align_backward_gather(go, [gf], rois, [ctx.spatial_scale], ctx.sample_num, ctx.variant)
            else:
                gf = torch.zeros((n, h, w, c), dtype=torch.float32, device=grad_output.device)
                F.rroi_align_backward(go, [gf], rois, [ctx.spatial_scale], ctx.sample_num, ctx.variant)
            grad_input = gf.permute(0, 3, 1, 2).to(grad_output.dtype)
        return grad_input


class RoIAlignFunction(_RoIAlignBase):
    """mmdet/ops/roi_align/roi_align.py:9-73 (aligned=False -> v1 legacy kernel, True -> v2)."""

    @staticmethod
    def forward(ctx, features, rois, out_size, spatial_scale, sample_num=0, aligned=True):
        assert rois.dim() == 2 and rois.size(1) == 5
        return _RoIAlignBase._forward(ctx, features, rois, out_size, spatial_scale, sample_num,
                                      _variant(aligned, legacy=not aligned))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _RoIAlignBase._backward(ctx, grad_output), None, None, None, None, None


class RoIAlignRotatedFunction(_RoIAlignBase):
    """Rotated RoIAlign: rois (K,6) [b,cx,cy,w,h,theta]; aligned=False keeps the legacy +1 box model."""

    @staticmethod
    def forward(ctx, features, rois, out_size, spatial_scale, sample_num=0, aligned=True):
        assert rois.dim() == 2 and rois.size(1) == 6
        return _RoIAlignBase._forward(ctx, features, rois, out_size, spatial_scale, sample_num,
                                      _variant(aligned, legacy=not aligned))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _RoIAlignBase._backward(ctx, grad_output), None, None, None, None, None


roi_align = RoIAlignFunction.apply
roi_align_rotated = RoIAlignRotatedFunction.apply


class RoIAlign(nn.Module):

    def __init__(self, out_size, spatial_scale, sample_num=0, use_torchvision=False, aligned=False):
        super(RoIAlign, self).__init__()
        self.out_size = _pair(out_size)
        self.spatial_scale = float(spatial_scale)
        self.aligned = aligned
        self.sample_num = int(sample_num)
        self.use_torchvision = use_torchvision
        assert not (use_torchvision and aligned), 'Torchvision does not support aligned RoIAlgin'

    def forward(self, features, rois):
        """features: NCHW images; rois: Bx5 boxes, first column is the index into N, then xyxy."""
        assert rois.dim() == 2 and rois.size(1) == 5
        if self.use_torchvision:
            from torchvision.ops import roi_align as tv_roi_align
            return tv_roi_align(features, rois, self.out_size, self.spatial_scale, self.sample_num)
        return roi_align(features, rois, self.out_size, self.spatial_scale, self.sample_num, self.aligned)

    def __repr__(self):
        format_str = self.__class__.__name__
        format_str += '(out_size={}, spatial_scale={}, sample_num={}'.format(
            self.out_size, self.spatial_scale, self.sample_num)
        format_str += ', use_torchvision={}, aligned={})'.format(self.use_torchvision, self.aligned)
        return format_str


class RoIAlignRotated(nn.Module):

    def __init__(self, out_size, spatial_scale, sample_num=0, aligned=False):
        super(RoIAlignRotated, self).__init__()
        self.out_size = _pair(out_size)
        self.spatial_scale = float(spatial_scale)
        self.aligned = aligned
        self.sample_num = int(sample_num)

    def forward(self, features, rois):
        """features: NCHW; rois: Bx6 [batch index, cx, cy, w, h, theta (rad)]."""
        assert rois.dim() == 2 and rois.size(1) == 6
        return roi_align_rotated(features, rois, self.out_size, self.spatial_scale, self.sample_num, self.aligned)

    def __repr__(self):
        return '{}(out_size={}, spatial_scale={}, sample_num={}, aligned={})'.format(
            self.__class__.__name__, self.out_size, self.spatial_scale, self.sample_num, self.aligned)
