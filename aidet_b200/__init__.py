"""aidet_b200 -- B200 (sm_100a) kernels for AIDet's oriented-bounding-box hot path.

    aidet_b200.ops    mirrors mmdet.ops   (nms, thetaobb_nms, pointobb_nms, batched_rnms,
                                           RoIAlign, RoIAlignRotated, roi_align, roi_align_rotated)
    aidet_b200.core   mirrors mmdet.core  (rbbox_overlaps, multiclass_nms, multiclass_thetaobb_nms, ...)

`install_as_mmdet()` registers these under the reference's module names so existing
`getattr(nms_wrapper, cfg['type'])` / `getattr(ops, roi_layer['type'])` lookups
(mmdet/core/post_processing/rbbox_nms.py:26-27, mmdet/models/roi_extractors/single_level.py:47-49)
resolve to the B200 implementations.  See INTEGRATION.md.
"""
__version__ = '0.1.0'

from . import core, ops  # noqa: E402,F401


def install_as_mmdet(force=False):
    """Expose the ops under `mmdet.ops.*` names inside an existing mmdet install (or stand-alone)."""
    import sys
    import types
    # `ops.nms` / `ops.roi_align` the attributes are functions (as in mmdet.ops); take the modules
    nms_pkg = sys.modules[__name__ + '.ops.nms']
    nms_wrapper = sys.modules[__name__ + '.ops.nms.nms_wrapper']
    names = {
        'mmdet.ops.nms': nms_pkg,
        'mmdet.ops.nms.nms_wrapper': nms_wrapper,
        'mmdet.ops.roi_align': sys.modules[__name__ + '.ops.roi_align'],
    }
    if 'mmdet' in sys.modules and not force:
        mm_ops = sys.modules.get('mmdet.ops')
        if mm_ops is not None:
            for k in ops.__all__:
                setattr(mm_ops, k, getattr(ops, k))
            nw = sys.modules.get('mmdet.ops.nms.nms_wrapper')
            if nw is not None:
                for k in ('nms', 'thetaobb_nms', 'pointobb_nms', 'batched_rnms'):
                    setattr(nw, k, getattr(nms_wrapper, k))
        mm_core = sys.modules.get('mmdet.core')
        if mm_core is not None:
            for k in core.__all__:
                setattr(mm_core, k, getattr(core, k))
        return
    for pkg in ('mmdet',):
        if pkg not in sys.modules:
            sys.modules[pkg] = types.ModuleType(pkg)
    sys.modules['mmdet.ops'] = ops
    sys.modules['mmdet.core'] = core
    sys.modules['mmdet'].ops = ops
    sys.modules['mmdet'].core = core
    for k, v in names.items():
        sys.modules[k] = v
