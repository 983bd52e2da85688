"""Multi-class NMS drivers, one kernel pass over all classes.

Mirrors mmdet/core/post_processing/bbox_nms.py:6-76 (`multiclass_nms`) and
rbbox_nms.py:6-62,64-119 (`multiclass_nms_with_index`, `thetaobb_nms_by_bbox_nms`), and
provides `multiclass_thetaobb_nms`, the call the reference left commented out
(mmdet/models/bbox_heads/rbbox_head.py:294-295).  The per-class Python loop and the
class-offset trick are both replaced by group ids handed to the batched kernel.
"""
import torch

from ...ops import functional as F
from ...ops.nms import nms_wrapper


def _select(multi_boxes, multi_scores, score_thr, dim, class_major=True, score_factors=None):
    """Score filter on the RAW scores -> (boxes (m,dim), scores (m,), labels (m,)).  class_major=True orders the
    candidates like the per-class loops of rbbox_nms.py:29-49; False like the boolean indexing of bbox_nms.py:41-47
    (box by box).  `score_factors` (n,) multiply the scores AFTER the filter, as bbox_nms.py:43-45 does."""
    num_classes = multi_scores.size(1) - 1
    if multi_boxes.shape[1] > dim:
        boxes = multi_boxes.view(multi_scores.size(0), -1, dim)[:, 1:]
    else:
        boxes = multi_boxes[:, None].expand(-1, num_classes, dim)
    scores = multi_scores[:, 1:]
    valid = scores > score_thr
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    if class_major:
        labels, rows = valid.t().nonzero(as_tuple=True)
    else:
        rows, labels = valid.nonzero(as_tuple=True)
    return boxes[rows, labels], scores[rows, labels], labels


def _finish(dets, labels, max_num):
    """bbox_nms.py:69-76 / rbbox_nms.py:52-57, as written there: `if k > max_num: sort by score, [:max_num]`.  With the
    default max_num = -1 the test is always true, so the reference returns the detections sorted by score WITHOUT THE
    LAST ONE (`inds[:-1]`); every config passes a positive max_per_img, and the quirk is reproduced, not corrected
    (tests/golden/golden_postproc_v1.npz holds the reference's outputs for -1, 40 and 100000)."""
    if dets.shape[0] > max_num:
        _, inds = dets[:, -1].sort(descending=True)
        inds = inds[:max_num]
        dets, labels = dets[inds], labels[inds]
    return dets, labels


def _multiclass(multi_boxes, multi_scores, score_thr, iou_thr, max_num, dim, plus_one, class_major=True,
                score_factors=None):
    boxes, scores, labels = _select(multi_boxes, multi_scores, score_thr, dim, class_major, score_factors)
    if boxes.numel() == 0:
        return multi_boxes.new_zeros((0, dim + 1)), multi_boxes.new_zeros((0, ), dtype=torch.long)
    num_classes = multi_scores.size(1) - 1
    keep = F.nms_batched(boxes, scores, labels, iou_thr, n_groups=num_classes, cmp_ge=False, plus_one=plus_one)
    dets = torch.cat([boxes[keep], scores[keep, None]], dim=1)
    return _finish(dets, labels[keep], max_num)


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """mmdet/core/post_processing/bbox_nms.py:6-76.  Returns (bboxes (k,5), labels (k,)), labels 0-based, in the
    reference's order (candidates box by box).  The class-offset trick of :57-59 becomes group ids of the batched
    kernel; `score_factors` scale the scores handed to the NMS and returned, the filter sees the raw scores."""
    cfg = nms_cfg.copy()
    nms_type = cfg.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError('multiclass_nms supports type="nms" only (got %r)' % nms_type)
    return _multiclass(multi_bboxes, multi_scores, score_thr, cfg.get('iou_thr', 0.5), max_num, 4, plus_one=True,
                       class_major=False, score_factors=score_factors)


def multiclass_thetaobb_nms(multi_rbboxes, multi_scores, score_thr, polygon_nms_iou_thr, max_num=-1,
                            out_dim_reg=5):
    """Rotated multi-class NMS (rbbox_head.py:294-295 call shape).

    multi_rbboxes (n, C*d) or (n, d), d = out_dim_reg (5 theta-OBB, 8 point-OBB);
    multi_scores (n, C+1) with the background in column 0.
    Returns (rbboxes (k, d+1), labels (k,)), labels 0-based, class-major then top-`max_num` by score.
    """
    return _multiclass(multi_rbboxes, multi_scores, score_thr, polygon_nms_iou_thr, max_num, out_dim_reg,
                       plus_one=False)


def multiclass_nms_with_index(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1):
    """mmdet/core/post_processing/rbbox_nms.py:6-62: also returns per-class masks and keep indices."""
    cfg = nms_cfg.copy()
    nms_type = cfg.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError('multiclass_nms_with_index supports type="nms" only (got %r)' % nms_type)
    num_classes = multi_scores.shape[1]
    valid = multi_scores[:, 1:] > score_thr
    bbox_cls_inds = [valid[:, i] for i in range(num_classes - 1)]
    boxes, scores, labels = _select(multi_bboxes, multi_scores, score_thr, 4)
    if boxes.numel() == 0:
        return (multi_bboxes.new_zeros((0, 5)), multi_bboxes.new_zeros((0, ), dtype=torch.long), bbox_cls_inds, [])
    keep = F.nms_batched(boxes, scores, labels, cfg.get('iou_thr', 0.5), n_groups=num_classes - 1, cmp_ge=False,
                         plus_one=True)
    # per-class keep indices, relative to that class's filtered subset (what nms_op returned at rbbox_nms.py:43-44)
    counts = torch.bincount(labels, minlength=num_classes - 1)
    starts = torch.cumsum(counts, 0) - counts
    kept_labels = labels[keep]
    bbox_keep_inds = []
    for c in range(num_classes - 1):
        if counts[c] == 0:
            continue
        bbox_keep_inds.append(keep[kept_labels == c] - starts[c])
    dets = torch.cat([boxes[keep], scores[keep, None]], dim=1)
    dets, out_labels = _finish(dets, kept_labels, max_num)
    return dets, out_labels, bbox_cls_inds, bbox_keep_inds


def thetaobb_nms_by_bbox_nms(multi_bboxes, multi_scores, bbox_cls_inds, bbox_keep_inds, max_num=-1, out_dim_reg=5):
    """mmdet/core/post_processing/rbbox_nms.py:64-119: gather OBBs with the keep indices of the HBB NMS (the slot
    where the reference left its rotated NMS commented out, :97-98; `multiclass_thetaobb_nms` is that call).  The
    reference pops `bbox_keep_inds` empty (:99); the caller's list is left alone here."""
    num_classes = multi_scores.shape[1]
    bboxes, labels = [], []
    keep_iter = iter(bbox_keep_inds)
    for i in range(1, num_classes):
        cls_inds = bbox_cls_inds[i - 1]
        if not cls_inds.any():
            continue
        if multi_bboxes.shape[1] == out_dim_reg:
            _bboxes = multi_bboxes[cls_inds, :]
        else:
            _bboxes = multi_bboxes[cls_inds, i * out_dim_reg:(i + 1) * out_dim_reg]
        cls_dets = torch.cat([_bboxes, multi_scores[cls_inds, i][:, None]], dim=1)
        cls_dets = cls_dets[next(keep_iter), :]
        bboxes.append(cls_dets)
        labels.append(multi_bboxes.new_full((cls_dets.shape[0], ), i - 1, dtype=torch.long))
    if bboxes:
        return _finish(torch.cat(bboxes), torch.cat(labels), max_num)
    return multi_bboxes.new_zeros((0, out_dim_reg + 1)), multi_bboxes.new_zeros((0, ), dtype=torch.long)


__all__ = ['multiclass_nms', 'multiclass_thetaobb_nms', 'multiclass_nms_with_index', 'thetaobb_nms_by_bbox_nms',
           'nms_wrapper']


def get_det_rbboxes(rois, cls_score, rbbox_pred, img_shape, scale_factor, rescale=False, cfg=None,
                    target_means=(0., 0., 0., 0., 0.), target_stds=(0.1, 0.1, 0.2, 0.2, 0.1), encode='thetaobb'):
    """`RBBoxHead.get_det_rbboxes_parallel` (mmdet/models/bbox_heads/rbbox_head.py:253-296) with the rotated NMS
    the reference left commented out (:294-295) in place of the HBB keep-index reuse (:288-293):
    softmax -> decode (delta2thetaobb | delta2pointobb, core/rbbox/transforms.py:356-395,458-505) -> rescale ->
    ONE batched rotated-NMS launch over all classes -> top `max_per_img`.  Everything stays on the device.

    rois (n,5) [batch, x1, y1, x2, y2]; cls_score (n, C+1) logits (or a list to average, :271-272); rbbox_pred
    (n, C*d) or None; cfg: object/dict with score_thr, polygon_nms_iou_thr (default 0.5, the dead config key of
    configs/dota/dota_v002_theta_obb_r50_v1_train.py:130) and max_per_img, or None to get (rbboxes, scores) back.
    """
    import torch.nn.functional as TF
    from ..rbbox import (delta2hobb, delta2pointobb, delta2thetaobb, hobb2pointobb, hobb_rescale, pointobb_rescale,
                         thetaobb_rescale)
    decode = {'thetaobb': delta2thetaobb, 'pointobb': delta2pointobb, 'hobb': delta2hobb}
    rescale_fn = {'thetaobb': thetaobb_rescale, 'pointobb': pointobb_rescale, 'hobb': hobb_rescale}
    if encode not in decode:
        raise ValueError("unknown encode %r (thetaobb | pointobb | hobb)" % (encode,))
    dim = 8 if encode == 'pointobb' else 5
    if isinstance(cls_score, list):
        cls_score = sum(cls_score) / float(len(cls_score))
    scores = TF.softmax(cls_score, dim=1) if cls_score is not None else None
    if rbbox_pred is not None:
        if len(target_means) != dim or len(target_stds) != dim:
            raise ValueError("encode=%r decodes %d deltas per box: target_means / target_stds must have %d entries, got %d / %d"
                             % (encode, dim, dim, len(target_means), len(target_stds)))
        means, stds = tuple(target_means), tuple(target_stds)
        rbboxes = decode[encode](rois[:, 1:], rbbox_pred, means, stds, img_shape)
    else:
        rbboxes = rois[:, 1:]
    if rescale:
        rbboxes = rescale_fn[encode](rbboxes.clone(), scale_factor, reverse_flag=True)
    if cfg is None:
        return rbboxes, scores
    get = cfg.get if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    if encode == 'hobb':        # the NMS kernels take theta-OBBs or corner lists: H-OBB -> 8 points (transforms.py:137-163)
        n = rbboxes.size(0)
        rbboxes = hobb2pointobb(rbboxes.reshape(n, -1, 5)).reshape(n, -1)
        dim = 8
    return multiclass_thetaobb_nms(rbboxes, scores, get('score_thr', 0.05), get('polygon_nms_iou_thr', 0.5),
                                   get('max_per_img', -1), out_dim_reg=dim)
