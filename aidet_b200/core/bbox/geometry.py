"""`bbox_overlaps` (mmdet/core/bbox/geometry.py:4-88) and its rotated counterpart `rbbox_overlaps`."""
import torch

from ...ops import functional as F


def rbbox_overlaps(rbboxes1, rbboxes2, mode='iou', is_aligned=False):
    """Overlap between two sets of oriented boxes.

    Args:
        rbboxes1 (Tensor): (m, 5) <cx, cy, w, h, theta[rad]> or (m, 8) <x1, y1, ..., x4, y4>.
        rbboxes2 (Tensor): (n, 5) or (n, 8); if is_aligned, m must equal n.
        mode (str): "iou" or "iof" (intersection over the area of rbboxes1).

    Returns:
        Tensor: (m, n), or (m,) if is_aligned.  True polygon areas (no `+1`, unlike
        the axis-aligned bbox_overlaps).  Empty inputs give the shapes of geometry.py:54-55.
    """
    assert mode in ['iou', 'iof']
    rows = rbboxes1.size(0)
    cols = rbboxes2.size(0)
    if is_aligned:
        assert rows == cols
    if rows * cols == 0:
        return rbboxes1.new(rows, 1) if is_aligned else rbboxes1.new(rows, cols)
    assert rbboxes1.size(-1) == rbboxes2.size(-1) and rbboxes1.size(-1) in (5, 8)
    if not rbboxes1.is_cuda:
        raise NotImplementedError('rbbox_overlaps has no CPU implementation')
    if is_aligned:
        out = F.riou_aligned(rbboxes1, rbboxes2, mode)
    else:
        out = F.riou_matrix(rbboxes1, rbboxes2, mode)
    return out.to(rbboxes1.dtype)


def bbox_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False):
    """Overlap between two sets of axis-aligned boxes (mmdet/core/bbox/geometry.py:4-88, legacy `+1` sides).

    bboxes1 (m, 4), bboxes2 (n, 4) in <x1, y1, x2, y2>; mode "iou" or "iof" (over the area of bboxes1).
    Returns (m, n), or (m,) if is_aligned; empty inputs give the shapes of geometry.py:54-55.  The reference
    materialises three (m, n, 2) temporaries with broadcasting; here it is one pass of the tiled kernel (fmt 4).
    """
    assert mode in ['iou', 'iof']
    rows = bboxes1.size(0)
    cols = bboxes2.size(0)
    if is_aligned:
        assert rows == cols
    if rows * cols == 0:
        return bboxes1.new(rows, 1) if is_aligned else bboxes1.new(rows, cols)
    if not bboxes1.is_cuda:
        raise NotImplementedError('bbox_overlaps has no CPU implementation here')
    b1, b2 = bboxes1[:, :4], bboxes2[:, :4]
    out = F.riou_aligned(b1, b2, mode) if is_aligned else F.riou_matrix(b1, b2, mode)
    return out.to(bboxes1.dtype)
