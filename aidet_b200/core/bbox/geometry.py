"""`rbbox_overlaps`: the rotated counterpart of mmdet/core/bbox/geometry.py:4-88 (`bbox_overlaps`)."""
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
