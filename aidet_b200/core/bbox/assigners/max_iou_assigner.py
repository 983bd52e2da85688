"""`MaxIoUAssigner` (mmdet/core/bbox/assigners/max_iou_assigner.py:8-195) for axis-aligned AND oriented boxes.

Same constructor, same `assign` / `assign_wrt_overlaps` methods, same AssignResult.  The box format follows the last
dimension of `gt_bboxes`: 4 = <x1,y1,x2,y2> (legacy `+1` overlaps, i.e. the reference's own behaviour), 5 =
<cx,cy,w,h,theta>, 8 = <x1,y1,...,x4,y4> (polygon overlaps, the rotated counterpart SURVEY 8f asks for).  The reference
builds the (k, n) overlap matrix, reduces it along both axes and then loops over the k truths in Python
(:155-182); here `assign` is one fused pass on the device that never writes the matrix (csrc/riou_assign.cu).
"""
import torch

from ....ops import functional as F
from .assign_result import AssignResult


class MaxIoUAssigner(object):
    """Assign a truth box or background to each box: -1 don't care, 0 negative, i > 0 the 1-based truth index.

    Args: see mmdet/core/bbox/assigners/max_iou_assigner.py:19-35 (pos_iou_thr, neg_iou_thr (float or pair),
    min_pos_iou, gt_max_assign_all, ignore_iof_thr, ignore_wrt_candidates, gpu_assign_thr).  `gpu_assign_thr` is
    accepted for config compatibility; there is no CPU path to fall back to, the fused kernel holds no (k, n) matrix.
    """

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1,
                 ignore_wrt_candidates=True, gpu_assign_thr=-1):
        self.pos_iou_thr = pos_iou_thr
        self.neg_iou_thr = neg_iou_thr
        self.min_pos_iou = min_pos_iou
        self.gt_max_assign_all = gt_max_assign_all
        self.ignore_iof_thr = ignore_iof_thr
        self.ignore_wrt_candidates = ignore_wrt_candidates
        self.gpu_assign_thr = gpu_assign_thr

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        """bboxes (n, fmt [+ score]), gt_bboxes (k, fmt) -> AssignResult (max_iou_assigner.py:52-120).

        >>> self = MaxIoUAssigner(0.5, 0.5)
        >>> bboxes = torch.Tensor([[0, 0, 10, 10], [10, 10, 20, 20]]).cuda()
        >>> gt_bboxes = torch.Tensor([[0, 0, 10, 9]]).cuda()
        >>> assert self.assign(bboxes, gt_bboxes).gt_inds.tolist() == [1, 0]
        """
        if gt_bboxes.numel() == 0:               # also the shape-(0,) tensor of tests/test_assigner.py:82
            return self._empty(0, bboxes.size(0), bboxes, gt_labels)
        fmt = gt_bboxes.size(-1)
        assert fmt in (4, 5, 8), 'gt_bboxes must be (k, 4), (k, 5) or (k, 8)'
        bboxes = bboxes[:, :fmt]
        k, n = gt_bboxes.size(0), bboxes.size(0)
        if n == 0:
            return self._empty(k, n, bboxes, gt_labels)
        if not bboxes.is_cuda:
            raise NotImplementedError('MaxIoUAssigner has no CPU implementation here')
        gt_inds, max_overlaps, labels = F.max_iou_assign(
            gt_bboxes, bboxes, self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou, self.gt_max_assign_all,
            gt_ignore=gt_bboxes_ignore, ignore_iof_thr=self.ignore_iof_thr,
            ignore_wrt_candidates=self.ignore_wrt_candidates, gt_labels=gt_labels)
        return AssignResult(k, gt_inds, max_overlaps.to(bboxes.dtype), labels=labels)

    def assign_wrt_overlaps(self, overlaps, gt_labels=None):
        """overlaps (k, n) between k truths and n boxes (-1 = ignored entry) -> AssignResult (:122-195)."""
        k, n = overlaps.size(0), overlaps.size(1)
        if k == 0 or n == 0:
            return self._empty(k, n, overlaps, gt_labels)
        if not overlaps.is_cuda:
            raise NotImplementedError('MaxIoUAssigner has no CPU implementation here')
        gt_inds, max_overlaps, labels = F.assign_wrt_overlaps(overlaps, self.pos_iou_thr, self.neg_iou_thr,
                                                              self.min_pos_iou, self.gt_max_assign_all, gt_labels)
        return AssignResult(k, gt_inds, max_overlaps.to(overlaps.dtype), labels=labels)

    @staticmethod
    def _empty(num_gts, num_bboxes, like, gt_labels):
        # max_iou_assigner.py:141-153: no truth -> everything background; no boxes -> empty
        gt_inds = like.new_full((num_bboxes, ), -1, dtype=torch.long)
        if num_gts == 0:
            gt_inds[:] = 0
        labels = None if gt_labels is None else like.new_zeros((num_bboxes, ), dtype=torch.long)
        return AssignResult(num_gts, gt_inds, like.new_zeros((num_bboxes, )), labels=labels)
