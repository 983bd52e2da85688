"""`ApproxMaxIoUAssigner` (mmdet/core/bbox/assigners/approx_max_iou_assigner.py:7-139): every square (base anchor) is
represented by `approxs_per_octave` approximations; a square's overlap with a truth is the best overlap of its
approximations, then the MaxIoUAssigner steps run on that (k, n) matrix.  Boxes may be HBB (n, 4), theta-OBB (n, 5) or
point-OBB (n, 8); the overlap matrix, the ignore IoF and the assignment steps are device kernels, the max over the
approximations is one tensor reduction."""
import torch

from ..geometry import bbox_overlaps, rbbox_overlaps
from .max_iou_assigner import MaxIoUAssigner


class ApproxMaxIoUAssigner(MaxIoUAssigner):

    def assign(self, approxs, squares, approxs_per_octave, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        """approxs (approxs_per_octave * n, fmt), squares (n, fmt), gt_bboxes (k, fmt) -> AssignResult (:52-139)."""
        num_squares = squares.size(0)
        num_gts = gt_bboxes.size(0) if gt_bboxes.numel() else 0
        if num_squares == 0 or num_gts == 0:
            return self._empty(num_gts, num_squares, squares, gt_labels)
        fmt = gt_bboxes.size(-1)
        assert fmt in (4, 5, 8), 'gt_bboxes must be (k, 4), (k, 5) or (k, 8)'
        if not squares.is_cuda:
            raise NotImplementedError('ApproxMaxIoUAssigner has no CPU implementation here')
        over = bbox_overlaps if fmt == 4 else rbbox_overlaps
        # re-organise to approxs_per_octave x num_squares (:99-101)
        approxs = torch.transpose(approxs[:, :fmt].reshape(num_squares, approxs_per_octave, fmt), 0, 1).contiguous().view(-1, fmt)
        all_overlaps = over(approxs, gt_bboxes)
        overlaps, _ = all_overlaps.view(approxs_per_octave, num_squares, num_gts).max(dim=0)
        overlaps = torch.transpose(overlaps, 0, 1).contiguous()
        bboxes = squares[:, :fmt]
        if (self.ignore_iof_thr > 0 and gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0
                and bboxes.numel() > 0):                                       # :119-129
            if self.ignore_wrt_candidates:
                ignore_max_overlaps, _ = over(bboxes, gt_bboxes_ignore, mode='iof').max(dim=1)
            else:
                ignore_max_overlaps, _ = over(gt_bboxes_ignore, bboxes, mode='iof').max(dim=0)
            overlaps[:, ignore_max_overlaps > self.ignore_iof_thr] = -1
        return self.assign_wrt_overlaps(overlaps, gt_labels)
