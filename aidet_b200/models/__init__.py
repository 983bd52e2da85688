"""`mmdet.models`-shaped namespace: only the callers that sit directly on the OBB hot path."""
from .losses import RotatedIoULoss, riou_loss, rotated_iou
from .roi_extractors import SingleRoIExtractor

__all__ = ['SingleRoIExtractor', 'RotatedIoULoss', 'riou_loss', 'rotated_iou']
