"""Regression / classification targets of the OBB heads (mmdet/core/rbbox/rbbox_target.py:8-88).

Per image: the sampled positives take the label and the encoded delta of the truth they were assigned to, weights 1;
negatives keep label 0 / zero targets with label weight 1.  The reference walks the positives in a Python loop
(`.cpu().numpy()`, `.tolist()`, `np.stack`, back to the device, `:59-76`); here the gather is one indexing op per
image and nothing leaves the device, so the step sits between the fused max-IoU assignment and the RoI extractor
without a host round trip.  Values are pinned to the reference function by tests/golden/golden_rbbox_v1.npz.
"""
import torch

from .transforms import hobb2delta, pointobb2delta, thetaobb2delta

_ENCODERS = {"thetaobb": thetaobb2delta, "pointobb": pointobb2delta, "hobb": hobb2delta}


def _encode_name(cfg):
    return cfg["encode"] if isinstance(cfg, dict) else cfg.encode


def rbbox_target_single(pos_proposals, neg_proposals, pos_assigned_gt_inds, gt_rbboxes, gt_labels, rbbox_test_cfg,
                        target_means, target_stds, out_dim_reg=5):
    """One image (rbbox_target.py:38-88) -> (labels (n,) long, label_weights (n,), rbbox_targets (n, d), rbbox_weights
    (n, d)) with the positives first."""
    num_pos, num_neg = pos_proposals.size(0), neg_proposals.size(0)
    n = num_pos + num_neg
    labels = pos_proposals.new_zeros(n, dtype=torch.long)
    label_weights = pos_proposals.new_zeros(n)
    targets = pos_proposals.new_zeros(n, out_dim_reg)
    weights = pos_proposals.new_zeros(n, out_dim_reg)
    if num_pos > 0:
        inds = pos_assigned_gt_inds.to(device=gt_rbboxes.device, dtype=torch.long)
        pos_gt = gt_rbboxes[inds].float().to(pos_proposals.device)
        targets[:num_pos] = _ENCODERS[_encode_name(rbbox_test_cfg)](pos_proposals, pos_gt, target_means, target_stds)
        labels[:num_pos] = gt_labels[inds].to(device=labels.device, dtype=torch.long)
        label_weights[:num_pos] = 1.0
        weights[:num_pos] = 1.0
    if num_neg > 0:
        label_weights[num_pos:] = 1.0
    return labels, label_weights, targets, weights


def rbbox_target(pos_proposals_list, neg_proposals_list, pos_assigned_gt_inds_list, gt_rbboxes_list, gt_labels_list,
                 rbbox_test_cfg, target_means, target_stds, out_dim_reg=5, concat=True):
    """All images of a batch (rbbox_target.py:8-35): four lists, or four concatenated tensors with concat=True."""
    per_image = [rbbox_target_single(p, n_, i, g, l, rbbox_test_cfg, target_means, target_stds, out_dim_reg)
                 for p, n_, i, g, l in zip(pos_proposals_list, neg_proposals_list, pos_assigned_gt_inds_list,
                                           gt_rbboxes_list, gt_labels_list)]
    cols = [list(c) for c in zip(*per_image)]
    if concat:
        cols = [torch.cat(c, 0) for c in cols]
    return tuple(cols)
