"""`mmdet.models`-shaped namespace: only the callers that sit directly on the OBB hot path."""
from .anchor_heads import rpn_get_bboxes, rpn_get_bboxes_single
from .losses import RotatedIoULoss, riou_loss, rotated_iou
from .roi_extractors import SingleRoIExtractor

__all__ = ['SingleRoIExtractor', 'RotatedIoULoss', 'riou_loss', 'rotated_iou', 'rpn_get_bboxes', 'rpn_get_bboxes_single']
