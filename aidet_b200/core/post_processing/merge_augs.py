"""Test-time-augmentation merging (mmdet/core/post_processing/merge_augs.py:8-78): the proposal form runs `nms`."""
import torch

from ...ops.nms import nms_wrapper
from ..bbox.transforms import bbox_mapping_back


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def merge_aug_proposals(aug_proposals, img_metas, rpn_test_cfg):
    """Merge augmented proposals (merge_augs.py:8-43): map every (n, 5) proposal set back to the original image,
    one NMS over their union (`rpn_test_cfg.nms_thr`), the best `rpn_test_cfg.max_num` by score.

    The NMS is the batched device kernel (one group); nothing is copied to the host except the keep count."""
    recovered = []
    for proposals, img_info in zip(aug_proposals, img_metas):
        p = proposals.clone()
        p[:, :4] = bbox_mapping_back(p[:, :4], img_info['img_shape'], img_info['scale_factor'], img_info['flip'])
        recovered.append(p)
    aug = torch.cat(recovered, dim=0)
    merged, _ = nms_wrapper.nms(aug, _get(rpn_test_cfg, 'nms_thr'))
    scores = merged[:, 4]
    _, order = scores.sort(0, descending=True)
    num = min(_get(rpn_test_cfg, 'max_num'), merged.shape[0])
    return merged[order[:num], :]


def merge_aug_bboxes(aug_bboxes, aug_scores, img_metas, rcnn_test_cfg):
    """Average the detections of the augmented views after mapping them back (merge_augs.py:46-70)."""
    recovered = []
    for bboxes, img_info in zip(aug_bboxes, img_metas):
        m = img_info[0]
        recovered.append(bbox_mapping_back(bboxes, m['img_shape'], m['scale_factor'], m['flip']))
    bboxes = torch.stack(recovered).mean(dim=0)
    if aug_scores is None:
        return bboxes
    return bboxes, torch.stack(aug_scores).mean(dim=0)


def merge_aug_scores(aug_scores):
    """merge_augs.py:73-78 (tensor branch)."""
    return torch.mean(torch.stack(aug_scores), dim=0)
