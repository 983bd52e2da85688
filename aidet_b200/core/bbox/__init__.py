from .geometry import rbbox_overlaps

__all__ = ['rbbox_overlaps']
