"""Generates tests/golden/golden_postproc_v1.npz by running the REFERENCE's own multi-class NMS drivers on the CPU:
    mmdet/core/post_processing/bbox_nms.py   multiclass_nms            
    mmdet/core/post_processing/rbbox_nms.py  multiclass_nms_with_index (a6), thetaobb_nms_by_bbox_nms (a7)
through the reference's own dispatcher mmdet/ops/nms/nms_wrapper.py (a1) and its own nms_cpu.cpp, compiled unmodified
into oracle/_ref/nms_cpu_ref.so by oracle/build_ref.py (a2).  Every file is loaded from /root/reference where it lies;
`mmdet.ops.nms.nms_cuda` is an empty stub (CPU tensors never reach it).  Run once in the dev container; the .npz is
committed and travels to the GPU box.

The CPU kernel suppresses on `>=` and the CUDA one on `>` (nms_cpu.cpp:56 vs nms_kernel.cu:61); the generator checks
that no pair of the inputs sits within 1e-6 of the threshold, so both give the same answer and the GPU tests may
demand bit-equal keeps.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref, oracle as O  # noqa: E402

REF = "/root/reference"
build_ref.build()
nms_cpu = build_ref.load()
for name, path in (("mmdet", "mmdet"), ("mmdet.ops", "mmdet/ops"), ("mmdet.ops.nms", "mmdet/ops/nms"),
                   ("mmdet.core", "mmdet/core"), ("mmdet.core.post_processing", "mmdet/core/post_processing")):
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REF, path)]      # a bare namespace: the package's own __init__ is NOT executed
    sys.modules[name] = mod
sys.modules["mmdet.ops.nms.nms_cpu"] = nms_cpu
sys.modules["mmdet.ops.nms"].nms_cpu = nms_cpu
sys.modules["mmdet.ops.nms.nms_cuda"] = types.ModuleType("mmdet.ops.nms.nms_cuda")
sys.modules["mmdet.ops.nms"].nms_cuda = sys.modules["mmdet.ops.nms.nms_cuda"]
nms_wrapper = importlib.import_module("mmdet.ops.nms.nms_wrapper")
sys.modules["mmdet.ops.nms"].nms_wrapper = nms_wrapper
multiclass_nms = importlib.import_module("mmdet.core.post_processing.bbox_nms").multiclass_nms
R = importlib.import_module("mmdet.core.post_processing.rbbox_nms")

# torch >= 1.5: `valid_mask.nonzero()` (bbox_nms.py:44) still works, with a deprecation warning only

rng = np.random.default_rng(21)
out = {}
n, C = 400, 6                                    # 6 foreground classes + background


def hbb(n, side=300.0):
    ctr = rng.uniform(0, side, (n // 8, 2)).repeat(8, 0) + rng.normal(0, 6, (n, 2))      # proposals pile on objects
    wh = rng.uniform(10, 70, (n, 2))
    return np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)


def near_threshold(boxes, scores_c, thr):
    """number of same-class candidate pairs whose +1 IoU lies within 1e-6 of thr"""
    if boxes.shape[0] < 2:
        return 0
    ov = O.hbb_overlaps(boxes.astype(np.float64), boxes.astype(np.float64)) if hasattr(O, "hbb_overlaps") else None
    if ov is None:
        b = boxes.astype(np.float64)
        lt = np.maximum(b[:, None, :2], b[None, :, :2]); rb = np.minimum(b[:, None, 2:], b[None, :, 2:])
        wh = np.clip(rb - lt + 1, 0, None); inter = wh[..., 0] * wh[..., 1]
        area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        ov = inter / (area[:, None] + area[None, :] - inter)
    iu = np.triu_indices(boxes.shape[0], 1)
    return int((np.abs(ov[iu] - thr) < 1e-6).sum())


logits = rng.normal(0, 1.6, (n, C + 1))
scores = torch.softmax(torch.from_numpy(logits), 1).float()
scores[:, 3] = 0.0                               # a class without candidates: the loops `continue` (rbbox_nms.py:32-33)
shared = torch.from_numpy(hbb(n))                                           # (n, 4): class-agnostic regression
per_class = torch.from_numpy(np.concatenate([hbb(n) for _ in range(C + 1)], 1))           # (n, 4 (C+1))
factors = torch.from_numpy(rng.uniform(0.3, 1.0, n).astype(np.float32))      # centerness-style factors < 1
obb5 = torch.from_numpy(rng.uniform(-1, 1, (n, 5 * (C + 1))).astype(np.float32))
obb8 = torch.from_numpy(rng.uniform(-1, 1, (n, 8)).astype(np.float32))
out.update(scores=scores.numpy(), shared=shared.numpy(), per_class=per_class.numpy(), factors=factors.numpy(),
           obb5=obb5.numpy(), obb8=obb8.numpy())
score_thr, iou_thr = 0.05, 0.5
out["score_thr"], out["iou_thr"] = np.float32(score_thr), np.float32(iou_thr)
cfg = dict(type='nms', iou_thr=iou_thr)

near = 0
for c in range(1, C + 1):
    m = (scores[:, c] > score_thr).numpy()
    near += near_threshold(shared.numpy()[m], None, iou_thr)
    near += near_threshold(per_class.numpy()[m, 4 * c:4 * c + 4], None, iou_thr)
assert near == 0, "regenerate with another seed: %d pairs within 1e-6 of the threshold" % near

for tag, boxes in (("shared", shared), ("per_class", per_class)):
    for max_num in (-1, 40, 100000):
        for fac in (None, factors):
            d, l = multiclass_nms(boxes, scores, score_thr, cfg, max_num, fac)
            key = "mc_%s_%d_%s" % (tag, max_num, "f" if fac is not None else "n")
            out[key + "_dets"], out[key + "_labels"] = d.numpy(), l.numpy()
        d, l, cls_inds, keep_inds = R.multiclass_nms_with_index(boxes, scores, score_thr, cfg, max_num)
        key = "wi_%s_%d" % (tag, max_num)
        out[key + "_dets"], out[key + "_labels"] = d.numpy(), l.numpy()
        out[key + "_cls_inds"] = np.stack([c.numpy() for c in cls_inds])
        out[key + "_keep_sizes"] = np.array([k.numel() for k in keep_inds])
        out[key + "_keep_cat"] = np.concatenate([k.numpy() for k in keep_inds])
        # a7: the OBB gather with the HBB keep indices (the function pops the list: hand it a copy)
        for otag, ob, dim in (("obb5", obb5, 5), ("obb8", obb8, 8)):
            if otag == "obb8" and tag == "per_class":
                continue
            d2, l2 = R.thetaobb_nms_by_bbox_nms(ob, scores, cls_inds, list(keep_inds), max_num, out_dim_reg=dim)
            out["%s_%s_dets" % (key, otag)], out["%s_%s_labels" % (key, otag)] = d2.numpy(), l2.numpy()
# nothing passes the score threshold
d, l = multiclass_nms(shared, scores, 2.0, cfg, 10)
d2, l2, ci, ki = R.multiclass_nms_with_index(shared, scores, 2.0, cfg, 10)
d3, l3 = R.thetaobb_nms_by_bbox_nms(obb5, scores, ci, list(ki), 10)
out["empty_shapes"] = np.array([d.shape[1], d2.shape[1], d3.shape[1], len(ci), len(ki)])
np.savez_compressed(os.path.join(HERE, "golden_postproc_v1.npz"), **out)
print("wrote golden_postproc_v1.npz:", len(out), "arrays; kept", out["mc_shared_100000_n_dets"].shape[0], "of", n, "x", C)
