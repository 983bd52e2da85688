"""CPU: the sampler mirrors (mmdet/core/bbox/samplers/*.py, assign_sampling.py) -- the reference's own
tests/test_sampler.py:7-93 cases plus invariants on oriented boxes.  The assignment itself needs the device, so the
AssignResult of the non-empty cases comes from the oracle restatement of MaxIoUAssigner."""
import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.core import (AssignResult, MaxIoUAssigner, PseudoSampler, RandomSampler, assign_and_sample, bbox2roi,
                             build_assigner, build_sampler, rbbox2roi)
from oracle import oracle as O

BBOXES = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
GTS = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])


def oracle_assign(bboxes, gts, labels, **kw):
    gi, mo, lb = O.max_iou_assign(bboxes.numpy(), gts.numpy(), gt_labels=labels.numpy(), **kw)
    return AssignResult(len(gts), torch.from_numpy(gi), torch.from_numpy(mo).float(), torch.from_numpy(lb))


def test_random_sampler_reference_cases():
    labels = torch.LongTensor([1, 2])
    ar = oracle_assign(BBOXES, GTS, labels, pos_iou_thr=0.5, neg_iou_thr=0.5, ignore_iof_thr=0.5, ignore_wrt_candidates=False,
                       gt_bboxes_ignore=np.array([[30, 30, 40, 40]], np.float32))
    sampler = RandomSampler(num=10, pos_fraction=0.5, neg_pos_ub=-1, add_gt_as_proposals=True)
    res = sampler.sample(ar, BBOXES, GTS, labels)                                    # tests/test_sampler.py:7-40
    assert len(res.pos_bboxes) == len(res.pos_inds) and len(res.neg_bboxes) == len(res.neg_inds)
    assert res.pos_is_gt.sum() == 2 and res.num_gts == 2                            # the truths were added as proposals
    assert res.pos_gt_bboxes.shape == (len(res.pos_inds), 4) and res.pos_gt_labels.tolist()[:2] == [1, 2]
    assert torch.equal(res.pos_gt_bboxes[:2], GTS)
    a = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5, ignore_iof_thr=0.5, ignore_wrt_candidates=False)
    empty_gt, empty_lb = torch.empty(0, 4), torch.empty(0, ).long()
    res = sampler.sample(a.assign(BBOXES, empty_gt, gt_labels=empty_lb), BBOXES, empty_gt, empty_lb)   # :43-66
    assert len(res.pos_bboxes) == len(res.pos_inds) == 0 and len(res.neg_bboxes) == len(res.neg_inds) == 4
    empty_b = torch.empty(0, 4)
    res = sampler.sample(a.assign(empty_b, GTS, gt_labels=labels), empty_b, GTS, labels)                # :69-93
    assert len(res.pos_bboxes) == len(res.pos_inds) == 2 and len(res.neg_inds) == 0                    # the added truths
    with pytest.raises(ValueError):
        sampler.sample(oracle_assign(BBOXES, GTS, labels, pos_iou_thr=0.5, neg_iou_thr=0.5), BBOXES, GTS, None)


def test_sampling_on_oriented_boxes_feeds_the_roi_format():
    boxes, gts, _, labels = synth.assign_case(3000, 25, side=512, seed=8)
    proposals = torch.cat([boxes, torch.rand(3000, 1)], 1)                           # (n, 6): score column, as an RPN emits
    ar = oracle_assign(boxes, gts, labels, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5)
    g = torch.Generator().manual_seed(1)
    sampler = RandomSampler(num=512, pos_fraction=0.25, neg_pos_ub=-1, add_gt_as_proposals=True, generator=g)
    res = sampler.sample(ar, proposals, gts, labels)
    assert res.bboxes.shape == (512, 5)                                              # all 5 OBB columns kept, score dropped
    assert len(res.pos_inds) == 128 and len(res.neg_inds) == 384
    assert len(torch.unique(torch.cat([res.pos_inds, res.neg_inds]))) == 512         # no index drawn twice
    assert bool((ar.gt_inds[res.pos_inds] > 0).all()) and bool((ar.gt_inds[res.neg_inds] == 0).all())
    assert torch.equal(res.pos_gt_bboxes, gts[res.pos_assigned_gt_inds])
    assert torch.equal(res.pos_gt_labels, labels[res.pos_assigned_gt_inds])
    rois = rbbox2roi([res.bboxes, res.bboxes[:7]])                                   # rbbox_cnn.py:177 with OBB RoIs
    assert rois.shape == (519, 6) and rois[512:, 0].tolist() == [1.0] * 7
    # neg_pos_ub caps the negatives; PseudoSampler keeps everything
    few = RandomSampler(num=512, pos_fraction=0.25, neg_pos_ub=1, add_gt_as_proposals=False, generator=g)
    ar2 = oracle_assign(boxes, gts, labels, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5)
    r2 = few.sample(ar2, boxes, gts, labels)
    assert len(r2.neg_inds) <= max(1, len(r2.pos_inds))
    r3 = PseudoSampler().sample(ar2, boxes, gts)
    assert len(r3.pos_inds) == int((ar2.gt_inds > 0).sum()) and len(r3.neg_inds) == int((ar2.gt_inds == 0).sum())
    assert bbox2roi([BBOXES]).shape == (4, 5)


def test_builders_and_assign_and_sample_on_empty_inputs():
    cfg = dict(assigner=dict(type='MaxIoUAssigner', pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5, ignore_iof_thr=-1),
               sampler=dict(type='RandomSampler', num=512, pos_fraction=0.25, neg_pos_ub=-1, add_gt_as_proposals=True))
    a, s = build_assigner(cfg['assigner']), build_sampler(cfg['sampler'])            # configs/dota/*: train_cfg.rcnn
    assert isinstance(a, MaxIoUAssigner) and a.min_pos_iou == 0.5 and isinstance(s, RandomSampler) and s.num == 512
    assert build_assigner(a) is a and build_sampler(s) is s
    with pytest.raises(TypeError):
        build_sampler(3)
    ar, sr = assign_and_sample(BBOXES, torch.empty(0, 4), None, torch.empty(0).long(), cfg)
    assert ar.gt_inds.tolist() == [0, 0, 0, 0] and len(sr.neg_inds) == 4 and len(sr.pos_inds) == 0
