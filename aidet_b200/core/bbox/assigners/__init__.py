from .assign_result import AssignResult
from .max_iou_assigner import MaxIoUAssigner

__all__ = ['AssignResult', 'MaxIoUAssigner']
