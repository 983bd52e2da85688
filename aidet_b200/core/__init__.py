"""`mmdet.core`-shaped namespace for the OBB hot path (mmdet/core/__init__.py:1-9)."""
from .bbox import rbbox_overlaps
from .post_processing import (multiclass_nms, multiclass_nms_with_index, multiclass_thetaobb_nms,
                              thetaobb_nms_by_bbox_nms)

__all__ = ['rbbox_overlaps', 'multiclass_nms', 'multiclass_nms_with_index', 'multiclass_thetaobb_nms',
           'thetaobb_nms_by_bbox_nms']
