"""`mmdet.ops`-shaped namespace (mmdet/ops/__init__.py:11,14,21-31) for the OBB hot path."""
from .nms import batched_rnms, nms, pointobb_nms, soft_nms, thetaobb_nms
from .roi_align import RoIAlign, RoIAlignRotated, roi_align, roi_align_rotated

__all__ = ['nms', 'soft_nms', 'thetaobb_nms', 'pointobb_nms', 'batched_rnms', 'RoIAlign', 'roi_align',
           'RoIAlignRotated', 'roi_align_rotated']
