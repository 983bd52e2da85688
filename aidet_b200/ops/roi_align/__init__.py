from .roi_align import RoIAlign, RoIAlignRotated, roi_align, roi_align_rotated

__all__ = ['roi_align', 'RoIAlign', 'roi_align_rotated', 'RoIAlignRotated']
