from .nms_wrapper import batched_rnms, nms, pointobb_nms, soft_nms, thetaobb_nms

__all__ = ['nms', 'soft_nms', 'thetaobb_nms', 'pointobb_nms', 'batched_rnms']
