"""Generates tests/golden/golden_assign_v1.npz by running the REFERENCE's own MaxIoUAssigner
(/root/reference/mmdet/core/bbox/assigners/max_iou_assigner.py, loaded file by file -- `import mmdet` itself
fails here for want of mmcv) on the CPU.  Run once in the dev container; the .npz is committed and travels.

Cases: `assign` on HBB boxes (the only format the reference's assigner knows) with and without ignore boxes and
labels, and `assign_wrt_overlaps` on float32 overlap matrices that hold exact ties and -1 (ignored) columns.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
for name, path in (("mmdet", "mmdet"), ("mmdet.utils", "mmdet/utils"), ("mmdet.core", "mmdet/core"),
                   ("mmdet.core.bbox", "mmdet/core/bbox"), ("mmdet.core.bbox.assigners", "mmdet/core/bbox/assigners")):
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REF, path)]      # a bare namespace: the package's own __init__ is NOT executed
    sys.modules[name] = mod
sys.modules["mmdet.utils"].util_mixins = importlib.import_module("mmdet.utils.util_mixins")
MaxIoUAssigner = importlib.import_module("mmdet.core.bbox.assigners.max_iou_assigner").MaxIoUAssigner
bbox_overlaps = importlib.import_module("mmdet.core.bbox.geometry").bbox_overlaps

CONFIGS = [  # (pos, neg, min_pos, assign_all, ignore_iof_thr, wrt_candidates)
    (0.5, 0.5, 0.0, True, -1, True),            # rcnn assigner of configs/*: pos 0.5 neg 0.5 min_pos 0.5 is below
    (0.5, 0.5, 0.5, True, -1, True),
    (0.7, 0.3, 0.3, True, 0.5, True),           # rpn assigner (ignore_iof_thr from the DOTA configs: -1; 0.5 exercises it)
    (0.7, (0.1, 0.3), 0.3, False, 0.5, False),
    (0.6, 0.4, 0.2, False, -1, True),
]

out = {"configs": np.array([[c[0], c[1][0] if isinstance(c[1], tuple) else 0.0, c[1][1] if isinstance(c[1], tuple) else c[1],
                             c[2], float(c[3]), c[4], float(c[5]), float(isinstance(c[1], tuple))] for c in CONFIGS])}
rng = np.random.default_rng(5)


def boxes(n, side=200):
    xy = rng.uniform(0, side, (n, 2))
    wh = rng.uniform(8, 90, (n, 2))
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


for ci, (pos, neg, mp, allg, ign, wrt) in enumerate(CONFIGS):
    a = MaxIoUAssigner(pos, neg, mp, allg, ign, wrt)
    gts, bxs, ignb = boxes(23), boxes(700), boxes(5)
    bxs[:23] = gts + rng.normal(0, 2, gts.shape).astype(np.float32)       # some real positives
    bxs[40:44] = bxs[36:40]                                               # duplicate boxes -> exact ties for a gt's max
    labels = rng.integers(1, 16, 23)
    r = a.assign(torch.from_numpy(bxs), torch.from_numpy(gts), torch.from_numpy(ignb), torch.from_numpy(labels))
    out["c%d_gts" % ci], out["c%d_boxes" % ci], out["c%d_ign" % ci], out["c%d_gt_labels" % ci] = gts, bxs, ignb, labels
    out["c%d_gt_inds" % ci] = r.gt_inds.numpy()
    out["c%d_max_overlaps" % ci] = r.max_overlaps.numpy()
    out["c%d_labels" % ci] = r.labels.numpy()
    out["c%d_overlaps" % ci] = bbox_overlaps(torch.from_numpy(gts), torch.from_numpy(bxs)).numpy()
    # a quantised matrix: many exact ties along both axes, -1 columns, an all-zero gt row
    ov = np.round(rng.uniform(0, 1, (17, 300)) * 20) / 20
    ov[:, rng.choice(300, 30, replace=False)] = -1
    ov[5] = np.where(ov[5] >= 0, 0.0, -1.0)
    ov = ov.astype(np.float32)
    r2 = a.assign_wrt_overlaps(torch.from_numpy(ov.copy()), torch.from_numpy(labels[:17]))
    out["c%d_mat" % ci] = ov
    out["c%d_mat_gt_inds" % ci] = r2.gt_inds.numpy()
    out["c%d_mat_max_overlaps" % ci] = r2.max_overlaps.numpy()
    out["c%d_mat_labels" % ci] = r2.labels.numpy()

# the doctest of max_iou_assigner.py:78-84
r = MaxIoUAssigner(0.5, 0.5).assign(torch.Tensor([[0, 0, 10, 10], [10, 10, 20, 20]]), torch.Tensor([[0, 0, 10, 9]]))
assert r.gt_inds.tolist() == [1, 0]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_assign_v1.npz"), **out)
print("wrote golden_assign_v1.npz", len(out), "arrays;",
      {k: int((out[k] > 0).sum()) for k in out if k.endswith("gt_inds")})
