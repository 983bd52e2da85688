"""The multi-class NMS drivers against OUTPUTS OF THE REFERENCE'S OWN FUNCTIONS (tests/golden/golden_postproc_v1.npz,
made by tests/golden/make_golden_postproc.py: mmdet/core/post_processing/bbox_nms.py + rbbox_nms.py through the
reference's nms_wrapper.py and its nms_cpu.cpp compiled unmodified) -- SURVEY.md 8a rows a5, a6, a7.

Two legs with the same assertions:
  * CPU (`not gpu`): the HOST logic of the drivers (candidate order, score factors, per-class keep indices, top-k quirk)
    with the batched-NMS call replaced by the oracle -- the only place a CPU stands in for the kernel, and only inside
    this test;
  * GPU (`gpu`): the real path, boxes on the device, `aidet_nms_batched_f32` underneath.
The generator verified that no candidate pair lies within 1e-6 of the IoU threshold, so keeps are compared bit for bit.
"""
import os

import numpy as np
import pytest
import torch

from aidet_b200 import core
from aidet_b200.core.post_processing import rbbox_nms as drivers
from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_postproc_v1.npz"))
SCORE_THR, IOU_THR = float(G["score_thr"]), float(G["iou_thr"])
CFG = dict(type='nms', iou_thr=IOU_THR)


def _oracle_nms_batched(boxes, scores, group_ids=None, iou_thr=0.5, n_groups=None, cmp_ge=False, plus_one=False, sync=True):
    keep, _ = O.nms(boxes.numpy(), scores.numpy(), float(iou_thr), groups=None if group_ids is None else group_ids.numpy(),
                    cmp_ge=cmp_ge, plus_one=plus_one)
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


@pytest.fixture(params=["host-logic", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cuda":
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        return torch.device("cuda", 0)
    monkeypatch.setattr(drivers.F, "nms_batched", _oracle_nms_batched)
    return torch.device("cpu")


def t(name, dev):
    return torch.from_numpy(G[name]).to(dev)


def same(got, name):
    want = G[name]
    got = got.cpu().numpy()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.array_equal(got, want), "%s differs: max |d| = %g" % (name, np.abs(got.astype(np.float64) - want).max())


@pytest.mark.parametrize("tag", ["shared", "per_class"])
@pytest.mark.parametrize("max_num", [-1, 40, 100000])
def test_multiclass_nms(dev, tag, max_num):
    # a5: bbox_nms.py:6-76 incl. score_factors (< 1: candidates pass on the raw score, are ranked on the product)
    for fac in (None, "factors"):
        d, l = core.multiclass_nms(t(tag, dev), t("scores", dev), SCORE_THR, CFG, max_num,
                                   None if fac is None else t(fac, dev))
        key = "mc_%s_%d_%s" % (tag, max_num, "f" if fac else "n")
        same(d, key + "_dets")
        same(l, key + "_labels")


@pytest.mark.parametrize("tag", ["shared", "per_class"])
@pytest.mark.parametrize("max_num", [-1, 40, 100000])
def test_multiclass_nms_with_index_and_obb_gather(dev, tag, max_num):
    # a6: rbbox_nms.py:6-62; a7: rbbox_nms.py:64-119 (theta-OBB (n, 5 (C+1)) and class-agnostic 8-point (n, 8) inputs)
    d, l, cls_inds, keep_inds = core.multiclass_nms_with_index(t(tag, dev), t("scores", dev), SCORE_THR, CFG, max_num)
    key = "wi_%s_%d" % (tag, max_num)
    same(d, key + "_dets")
    same(l, key + "_labels")
    same(torch.stack(cls_inds), key + "_cls_inds")
    assert [k.numel() for k in keep_inds] == G[key + "_keep_sizes"].tolist()
    same(torch.cat(keep_inds), key + "_keep_cat")
    n_before = len(keep_inds)
    for otag, dim in (("obb5", 5), ("obb8", 8)):
        if otag == "obb8" and tag == "per_class":
            continue
        d2, l2 = core.thetaobb_nms_by_bbox_nms(t(otag, dev), t("scores", dev), cls_inds, keep_inds, max_num, out_dim_reg=dim)
        same(d2, "%s_%s_dets" % (key, otag))
        same(l2, "%s_%s_labels" % (key, otag))
    assert len(keep_inds) == n_before                     # the caller's list survives (the reference pops it empty)


def test_nothing_passes_the_score_threshold(dev):
    d, l = core.multiclass_nms(t("shared", dev), t("scores", dev), 2.0, CFG, 10)
    d2, l2, ci, ki = core.multiclass_nms_with_index(t("shared", dev), t("scores", dev), 2.0, CFG, 10)
    d3, l3 = core.thetaobb_nms_by_bbox_nms(t("obb5", dev), t("scores", dev), ci, ki, 10)
    assert [d.shape[1], d2.shape[1], d3.shape[1], len(ci), len(ki)] == G["empty_shapes"].tolist()
    assert d.shape[0] == d2.shape[0] == d3.shape[0] == 0 and l.dtype == l2.dtype == l3.dtype == torch.long
