"""`AssignResult` (mmdet/core/bbox/assigners/assign_result.py:6-192): the container MaxIoUAssigner returns."""
import torch


class AssignResult(object):
    """Assignments between predicted and truth boxes.

    Attributes:
        num_gts (int): number of truth boxes considered.
        gt_inds (LongTensor): per predicted box the 1-based index of its truth box, 0 = unassigned
            (background), -1 = ignore.
        max_overlaps (FloatTensor): per predicted box the largest overlap with any truth box.
        labels (None | LongTensor): per predicted box the category label of its truth box.
    """

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds
        self.max_overlaps = max_overlaps
        self.labels = labels

    @property
    def num_preds(self):
        return len(self.gt_inds)

    @property
    def info(self):
        return {'num_gts': self.num_gts, 'num_preds': self.num_preds, 'gt_inds': self.gt_inds,
                'max_overlaps': self.max_overlaps, 'labels': self.labels}

    def __repr__(self):
        def shp(t):
            return repr(t) if t is None else repr(tuple(t.shape))
        return '<AssignResult(num_gts=%r, gt_inds.shape=%s, max_overlaps.shape=%s, labels.shape=%s)>' % (
            self.num_gts, shp(self.gt_inds), shp(self.max_overlaps), shp(self.labels))

    def add_gt_(self, gt_labels):
        """Prepend the truth boxes themselves as proposals (assign_result.py:182-192)."""
        self_inds = torch.arange(1, len(gt_labels) + 1, dtype=torch.long, device=gt_labels.device)
        self.gt_inds = torch.cat([self_inds, self.gt_inds])
        self.max_overlaps = torch.cat([self.max_overlaps.new_ones(len(gt_labels)), self.max_overlaps])
        if self.labels is not None:
            self.labels = torch.cat([gt_labels, self.labels])
