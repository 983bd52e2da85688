"""`mmdet.core`-shaped namespace for the OBB hot path (mmdet/core/__init__.py:1-9)."""
from .bbox import (ApproxMaxIoUAssigner, AssignResult, MaxIoUAssigner, PseudoSampler, RandomSampler, SamplingResult,
                   assign_and_sample, build_assigner, build_sampler, bbox2delta, bbox2result, bbox2roi, bbox_flip,
                   bbox_mapping, bbox_mapping_back, bbox_overlaps, delta2bbox, rbbox2roi, rbbox_overlaps, roi2bbox)
from .post_processing import (get_det_rbboxes, merge_aug_bboxes, merge_aug_proposals, merge_aug_scores, multiclass_nms,
                              multiclass_nms_with_index, multiclass_thetaobb_nms, thetaobb_nms_by_bbox_nms)
from .rbbox import (delta2hobb, delta2pointobb, delta2thetaobb, hobb2delta, hobb2pointobb, hobb_rescale, pointobb2bbox,
                    pointobb2delta, pointobb_rescale, rbbox2result, thetaobb2delta, thetaobb2pointobb, thetaobb_rescale)

__all__ = ['bbox_overlaps', 'rbbox_overlaps', 'AssignResult', 'MaxIoUAssigner', 'ApproxMaxIoUAssigner', 'RandomSampler', 'PseudoSampler', 'SamplingResult',
           'assign_and_sample', 'build_assigner', 'build_sampler', 'bbox2delta', 'delta2bbox',
           'bbox_flip', 'bbox2roi', 'rbbox2roi', 'roi2bbox', 'bbox2result', 'bbox_mapping', 'bbox_mapping_back', 'merge_aug_proposals', 'merge_aug_bboxes',
           'merge_aug_scores', 'multiclass_nms', 'multiclass_nms_with_index', 'multiclass_thetaobb_nms',
           'thetaobb_nms_by_bbox_nms', 'get_det_rbboxes', 'delta2pointobb', 'delta2thetaobb', 'pointobb2bbox',
           'pointobb2delta', 'pointobb_rescale', 'rbbox2result', 'thetaobb2delta', 'thetaobb2pointobb',
           'thetaobb_rescale', 'delta2hobb', 'hobb2delta', 'hobb2pointobb', 'hobb_rescale']
