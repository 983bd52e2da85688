"""Rotated IoU loss: the oriented-box counterpart of `iou_loss` / `IoULoss`
(mmdet/models/losses/iou_loss.py:10-27,129-165) with the same call signature.

The reference's loss is `-log(bbox_overlaps(pred, target, is_aligned=True).clamp(min=eps))` and gets its gradient from
autograd through the broadcast min/max of bbox_overlaps.  Polygon clipping has no such autograd graph, so the aligned
rotated overlap is one autograd Function whose backward is a kernel (aidet_riou_aligned_grad_f32): the derivative of
the intersection area is the boundary integral of the normal velocity over the parts of each box's edges inside the
other box (csrc/geom.cuh: rect_inter_grad), exact wherever the overlap is differentiable.
"""
import functools

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ...ops import functional as F


class _RotatedIoU(Function):

    @staticmethod
    def forward(ctx, pred, target, mode):
        if not pred.is_cuda:
            raise NotImplementedError('rotated_iou has no CPU implementation')
        assert pred.shape == target.shape and pred.dim() == 2 and pred.size(1) in (5, 8), \
            'rotated_iou takes aligned (n, 5) <cx, cy, w, h, theta> or (n, 8) <x1, y1, ..., x4, y4> boxes'
        ctx.mode = mode
        ctx.save_for_backward(pred, target)
        return F.riou_aligned(pred, target, mode).to(pred.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_ov):
        pred, target = ctx.saved_tensors
        want_t = ctx.needs_input_grad[1]
        _, gp, gt = F.riou_aligned_grad(pred, target, grad_ov, ctx.mode, want_b=want_t)
        return (gp.to(pred.dtype) if ctx.needs_input_grad[0] else None, gt.to(target.dtype) if want_t else None, None)


def rotated_iou(pred, target, mode='iou'):
    """Differentiable aligned overlap of oriented boxes: pred, target (n, 5) <cx, cy, w, h, theta[rad]> or (n, 8) convex
    <x1, y1, ..., x4, y4> -> (n,)."""
    if pred.size(0) == 0:
        return pred.new_zeros((0, ))
    return _RotatedIoU.apply(pred, target, mode)


def _reduce(loss, weight, reduction, avg_factor):
    # mmdet/models/losses/utils.py:26-52 (weight_reduce_loss)
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        assert reduction == 'none'
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def _weighted(loss_func):
    # mmdet/models/losses/utils.py:55-98 (weighted_loss)
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction='mean', avg_factor=None, **kwargs):
        return _reduce(loss_func(pred, target, **kwargs), weight, reduction, avg_factor)
    return wrapper


@_weighted
def riou_loss(pred, target, eps=1e-6):
    """-log(rotated IoU) of aligned oriented boxes, pred / target (n, 5) or (n, 8) (iou_loss.py:10-27 for OBBs)."""
    ious = rotated_iou(pred, target).clamp(min=eps)
    return -ious.log()


class RotatedIoULoss(nn.Module):
    """`IoULoss` (iou_loss.py:129-165) on <cx, cy, w, h, theta> or 8-point boxes: same arguments, same reduction rules."""

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super(RotatedIoULoss, self).__init__()
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        if weight is not None and not torch.any(weight > 0):
            return (pred * weight).sum()  # 0
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * riou_loss(pred, target, weight, eps=self.eps, reduction=reduction,
                                            avg_factor=avg_factor, **kwargs)
