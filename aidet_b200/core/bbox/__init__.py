from .assign_sampling import assign_and_sample, build_assigner, build_sampler
from .assigners import ApproxMaxIoUAssigner, AssignResult, MaxIoUAssigner
from .geometry import bbox_overlaps, rbbox_overlaps
from .samplers import BaseSampler, PseudoSampler, RandomSampler, SamplingResult
from .transforms import (bbox2delta, bbox2result, bbox2roi, bbox_flip, bbox_mapping, bbox_mapping_back, delta2bbox,
                         rbbox2roi, roi2bbox)

__all__ = ['bbox_overlaps', 'rbbox_overlaps', 'AssignResult', 'MaxIoUAssigner', 'ApproxMaxIoUAssigner', 'bbox2delta', 'delta2bbox', 'bbox_flip',
           'bbox_mapping', 'bbox_mapping_back', 'bbox2roi', 'rbbox2roi', 'roi2bbox', 'bbox2result', 'assign_and_sample', 'build_assigner', 'build_sampler',
           'BaseSampler', 'PseudoSampler', 'RandomSampler', 'SamplingResult']
