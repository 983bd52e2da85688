"""The samplers that consume an AssignResult on the way to the RoI extractor
(mmdet/core/bbox/samplers/{base_sampler,random_sampler,pseudo_sampler,sampling_result}.py): assign -> sample ->
bbox2roi / rbbox2roi -> RoIAlign.  Plain tensor code on whatever device the boxes live on; the only change against the
reference is that oriented boxes keep all their columns (the reference cuts every box to 4 numbers,
base_sampler.py:58)."""
import torch


def _box_dim(bboxes, gt_bboxes):
    """4, 5 or 8 coordinate columns: taken from the truths when there are any, else from the candidates (whose last
    column may be a score)."""
    if gt_bboxes.dim() == 2 and gt_bboxes.size(-1) in (4, 5, 8):
        return gt_bboxes.size(-1)
    d = bboxes.size(-1)
    return d if d in (4, 8) else d - 1 if d in (5, 6, 9) and d - 1 in (4, 5, 8) else 4


class SamplingResult(object):
    """sampling_result.py:6-60: the sampled positives / negatives, their boxes, and the truths of the positives."""

    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
        self.pos_inds = pos_inds
        self.neg_inds = neg_inds
        self.pos_bboxes = bboxes[pos_inds]
        self.neg_bboxes = bboxes[neg_inds]
        self.pos_is_gt = gt_flags[pos_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        d = bboxes.size(-1)
        if gt_bboxes.numel() == 0:
            assert self.pos_assigned_gt_inds.numel() == 0
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, d)
        else:
            if gt_bboxes.dim() < 2:
                gt_bboxes = gt_bboxes.view(-1, d)
            self.pos_gt_bboxes = gt_bboxes[self.pos_assigned_gt_inds, :]
        self.pos_gt_labels = assign_result.labels[pos_inds] if assign_result.labels is not None else None

    @property
    def bboxes(self):
        return torch.cat([self.pos_bboxes, self.neg_bboxes])

    def to(self, device):
        for key, value in self.__dict__.items():
            if isinstance(value, torch.Tensor):
                self.__dict__[key] = value.to(device)
        return self

    def __repr__(self):
        return '<SamplingResult(num_gts=%d, pos=%d, neg=%d)>' % (self.num_gts, self.pos_inds.numel(), self.neg_inds.numel())


class BaseSampler(object):
    """base_sampler.py:8-98."""

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, **kwargs):
        self.num = num
        self.pos_fraction = pos_fraction
        self.neg_pos_ub = neg_pos_ub
        self.add_gt_as_proposals = add_gt_as_proposals
        self.pos_sampler = self
        self.neg_sampler = self

    def _sample_pos(self, assign_result, num_expected, **kwargs):
        raise NotImplementedError

    def _sample_neg(self, assign_result, num_expected, **kwargs):
        raise NotImplementedError

    def sample(self, assign_result, bboxes, gt_bboxes, gt_labels=None, **kwargs):
        """base_sampler.py:31-98: optionally prepend the truths as proposals, draw <= num * pos_fraction positives,
        fill up with negatives (capped at neg_pos_ub x positives)."""
        if bboxes.dim() < 2:
            bboxes = bboxes[None, :]
        bboxes = bboxes[:, :_box_dim(bboxes, gt_bboxes)]
        gt_flags = bboxes.new_zeros((bboxes.shape[0], ), dtype=torch.uint8)
        if self.add_gt_as_proposals and len(gt_bboxes) > 0:
            if gt_labels is None:
                raise ValueError('gt_labels must be given when add_gt_as_proposals is True')
            bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
            assign_result.add_gt_(gt_labels)
            gt_flags = torch.cat([bboxes.new_ones(gt_bboxes.shape[0], dtype=torch.uint8), gt_flags])
        num_expected_pos = int(self.num * self.pos_fraction)
        pos_inds = self.pos_sampler._sample_pos(assign_result, num_expected_pos, bboxes=bboxes, **kwargs).unique()
        num_sampled_pos = pos_inds.numel()
        num_expected_neg = self.num - num_sampled_pos
        if self.neg_pos_ub >= 0:
            num_expected_neg = min(num_expected_neg, int(self.neg_pos_ub * max(1, num_sampled_pos)))
        neg_inds = self.neg_sampler._sample_neg(assign_result, num_expected_neg, bboxes=bboxes, **kwargs).unique()
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)


class RandomSampler(BaseSampler):
    """random_sampler.py:6-63; `generator=` makes the draw reproducible (the reference keeps a numpy rng it never uses)."""

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, generator=None, **kwargs):
        super(RandomSampler, self).__init__(num, pos_fraction, neg_pos_ub, add_gt_as_proposals)
        self.generator = generator

    def random_choice(self, gallery, num):
        assert len(gallery) >= num
        if not isinstance(gallery, torch.Tensor):
            gallery = torch.as_tensor(gallery, dtype=torch.long)
        if self.generator is not None:          # generators live on one device: draw there, index where the data is
            perm = torch.randperm(gallery.numel(), generator=self.generator, device=self.generator.device)[:num]
            perm = perm.to(gallery.device)
        else:
            perm = torch.randperm(gallery.numel(), device=gallery.device)[:num]
        return gallery[perm]

    def _sample_pos(self, assign_result, num_expected, **kwargs):
        pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).flatten()
        return pos_inds if pos_inds.numel() <= num_expected else self.random_choice(pos_inds, num_expected)

    def _sample_neg(self, assign_result, num_expected, **kwargs):
        neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).flatten()
        return neg_inds if neg_inds.numel() <= num_expected else self.random_choice(neg_inds, num_expected)


class PseudoSampler(BaseSampler):
    """pseudo_sampler.py:7-26: keeps every positive and every negative."""

    def __init__(self, **kwargs):
        pass

    def sample(self, assign_result, bboxes, gt_bboxes, **kwargs):
        pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).flatten().unique()
        neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).flatten().unique()
        bboxes = bboxes[:, :_box_dim(bboxes, gt_bboxes)]
        gt_flags = bboxes.new_zeros(bboxes.shape[0], dtype=torch.uint8)
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)
