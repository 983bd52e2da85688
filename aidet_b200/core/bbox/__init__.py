from .assigners import AssignResult, MaxIoUAssigner
from .geometry import bbox_overlaps, rbbox_overlaps

__all__ = ['bbox_overlaps', 'rbbox_overlaps', 'AssignResult', 'MaxIoUAssigner']
