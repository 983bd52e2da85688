from .iou_loss import RotatedIoULoss, riou_loss, rotated_iou

__all__ = ['RotatedIoULoss', 'riou_loss', 'rotated_iou']
