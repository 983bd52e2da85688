from .assigners import ApproxMaxIoUAssigner, AssignResult, MaxIoUAssigner
from .geometry import bbox_overlaps, rbbox_overlaps
from .transforms import bbox2delta, bbox_flip, bbox_mapping, bbox_mapping_back, delta2bbox

__all__ = ['bbox_overlaps', 'rbbox_overlaps', 'AssignResult', 'MaxIoUAssigner', 'ApproxMaxIoUAssigner', 'bbox2delta', 'delta2bbox', 'bbox_flip',
           'bbox_mapping', 'bbox_mapping_back']
