"""`mmdet.models`-shaped namespace: only the callers that sit directly on the OBB hot path."""
from .roi_extractors import SingleRoIExtractor

__all__ = ['SingleRoIExtractor']
