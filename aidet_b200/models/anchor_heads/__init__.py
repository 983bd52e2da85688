from .rpn_head import rpn_get_bboxes, rpn_get_bboxes_single

__all__ = ['rpn_get_bboxes', 'rpn_get_bboxes_single']
