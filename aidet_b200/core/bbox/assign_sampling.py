"""`build_assigner` / `build_sampler` / `assign_and_sample` (mmdet/core/bbox/assign_sampling.py:6-33): the reference looks
the classes up by name with mmcv.runner.obj_from_dict; the same dict configs work here."""
from . import assigners, samplers


def _from_dict(cfg, namespace, base, **default_args):
    if isinstance(cfg, base):
        return cfg
    if not isinstance(cfg, dict):
        raise TypeError('Invalid type {} for building an assigner / sampler'.format(type(cfg)))
    args = dict(cfg)
    cls = getattr(namespace, args.pop('type'))
    for k, v in default_args.items():
        args.setdefault(k, v)
    return cls(**args)


def build_assigner(cfg, **kwargs):
    return _from_dict(cfg, assigners, assigners.MaxIoUAssigner, **kwargs)


def build_sampler(cfg, **kwargs):
    return _from_dict(cfg, samplers, samplers.BaseSampler, **kwargs)


def assign_and_sample(bboxes, gt_bboxes, gt_bboxes_ignore, gt_labels, cfg):
    """cfg.assigner / cfg.sampler (attributes or keys) -> (AssignResult, SamplingResult), assign_sampling.py:26-33."""
    get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    bbox_assigner = build_assigner(get('assigner'))
    bbox_sampler = build_sampler(get('sampler'))
    assign_result = bbox_assigner.assign(bboxes, gt_bboxes, gt_bboxes_ignore, gt_labels)
    sampling_result = bbox_sampler.sample(assign_result, bboxes, gt_bboxes, gt_labels)
    return assign_result, sampling_result
