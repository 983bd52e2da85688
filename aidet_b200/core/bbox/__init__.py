from .assigners import AssignResult, MaxIoUAssigner
from .geometry import bbox_overlaps, rbbox_overlaps
from .transforms import bbox2delta, delta2bbox

__all__ = ['bbox_overlaps', 'rbbox_overlaps', 'AssignResult', 'MaxIoUAssigner', 'bbox2delta', 'delta2bbox']
