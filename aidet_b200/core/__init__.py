"""`mmdet.core`-shaped namespace for the OBB hot path (mmdet/core/__init__.py:1-9)."""
from .bbox import (AssignResult, MaxIoUAssigner, bbox_overlaps, rbbox_overlaps)
from .post_processing import (get_det_rbboxes, multiclass_nms, multiclass_nms_with_index, multiclass_thetaobb_nms,
                              thetaobb_nms_by_bbox_nms)
from .rbbox import (rbbox_target, rbbox_target_single, delta2hobb, delta2pointobb, delta2thetaobb, hobb2delta,
                    hobb2pointobb, hobb_flip, hobb_mapping, hobb_mapping_back, hobb_rescale, pointobb2bbox,
                    pointobb2delta, pointobb2thetaobb, pointobb_best_point_sort, pointobb_extreme_sort, pointobb_flip,
                    pointobb_mapping, pointobb_mapping_back, pointobb_rescale, rbbox2result, rbbox2roi,
                    thetaobb2delta, thetaobb2hobb, thetaobb2pointobb, thetaobb_flip, thetaobb_mapping,
                    thetaobb_mapping_back, thetaobb_rescale)

__all__ = ['AssignResult', 'MaxIoUAssigner', 'bbox_overlaps', 'rbbox_overlaps', 'get_det_rbboxes', 'multiclass_nms',
           'multiclass_nms_with_index', 'multiclass_thetaobb_nms', 'thetaobb_nms_by_bbox_nms', 'rbbox_target',
           'rbbox_target_single', 'delta2hobb', 'delta2pointobb', 'delta2thetaobb', 'hobb2delta', 'hobb2pointobb',
           'hobb_flip', 'hobb_mapping', 'hobb_mapping_back', 'hobb_rescale', 'pointobb2bbox', 'pointobb2delta',
           'pointobb2thetaobb', 'pointobb_best_point_sort', 'pointobb_extreme_sort', 'pointobb_flip',
           'pointobb_mapping', 'pointobb_mapping_back', 'pointobb_rescale', 'rbbox2result', 'rbbox2roi',
           'thetaobb2delta', 'thetaobb2hobb', 'thetaobb2pointobb', 'thetaobb_flip', 'thetaobb_mapping',
           'thetaobb_mapping_back', 'thetaobb_rescale']
