"""Generates tests/golden/golden_v1.npz (run once in the dev container, where /root/reference,
cv2 and torchvision are available; the file it writes is committed and travels to the GPU box).

Sources of truth:
  theta_points_cv2   cv2.boxPoints, the codec the reference uses (mmdet/core/rbbox/transforms.py:45-55)
  riou_cv2           cv2.rotatedRectangleIntersection + contourArea (float32, loose)
  hbb_keep_ref_0p5   the reference's nms_cpu.cpp compiled unmodified (oracle/_ref)
  roi_fwd_* / bwd    torchvision.ops.roi_align on CPU in float64, the implementation the reference
                     itself offers (mmdet/ops/roi_align/roi_align.py:138-141); v1 = rois with x2+1,y2+1
"""
import os
import sys

import cv2
import numpy as np
import torch
from torchvision.ops import roi_align as tv_roi_align

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aidet_b200 import synth  # noqa: E402
from oracle import build_ref  # noqa: E402

out = {}
boxes, _ = synth.dota_boxes(48, side=300, seed=100)
b = boxes.numpy()
out["theta_boxes"] = b
pts = np.stack([cv2.boxPoints(((float(r[0]), float(r[1])), (float(r[2]), float(r[3])), float(r[4]) * 180 / np.pi)).reshape(-1)
                for r in b])
out["theta_points_cv2"] = pts.astype(np.float64)
m = np.zeros((48, 48))
for i in range(48):
    for j in range(48):
        r1 = ((float(b[i, 0]), float(b[i, 1])), (float(b[i, 2]), float(b[i, 3])), float(np.degrees(b[i, 4])))
        r2 = ((float(b[j, 0]), float(b[j, 1])), (float(b[j, 2]), float(b[j, 3])), float(np.degrees(b[j, 4])))
        ret, p = cv2.rotatedRectangleIntersection(r1, r2)
        ar = cv2.contourArea(cv2.convexHull(p)) if ret > 0 and p is not None and len(p) > 2 else 0.0
        m[i, j] = ar / (b[i, 2] * b[i, 3] + b[j, 2] * b[j, 3] - ar)
out["riou_cv2"] = m

rng = np.random.default_rng(7)
xy = rng.uniform(0, 200, (400, 2))
wh = rng.uniform(5, 80, (400, 2))
dets = np.concatenate([xy, xy + wh, rng.uniform(0, 1, (400, 1))], 1).astype(np.float32)
build_ref.build()
ref = build_ref.load()
out["hbb_dets"] = dets
out["hbb_keep_ref_0p5"] = ref.nms(torch.from_numpy(dets), 0.5).numpy()

g = torch.Generator().manual_seed(11)
feat = torch.randn(2, 16, 15, 15, generator=g)                   # mmdet/ops/roi_align/gradcheck.py:11-30 shapes
x1 = torch.rand(20, generator=g) * 90
y1 = torch.rand(20, generator=g) * 90
rois = torch.stack([torch.randint(0, 2, (20,), generator=g).float(), x1, y1, x1 + torch.rand(20, generator=g) * 60 + 8,
                    y1 + torch.rand(20, generator=g) * 60 + 8], 1)      # >= 1 feature px: where v1 == torchvision
out["feat_nchw"] = feat.numpy()
out["rois5"] = rois.numpy()
r1 = rois.double().clone()
r1[:, 3:] += 1
out["roi_fwd_v1"] = tv_roi_align(feat.double(), r1, (3, 3), 0.125, 2, aligned=False).numpy()
fd = feat.double().requires_grad_(True)
y = tv_roi_align(fd, rois.double(), (3, 3), 0.125, 2, aligned=True)
out["roi_fwd_v2a"] = y.detach().numpy()
go = torch.randn(y.shape, generator=g)
y.backward(go.double())
out["roi_grad_out"] = go.numpy()
out["roi_bwd_v2a"] = fd.grad.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
print("wrote golden_v1.npz", {k: v.shape for k, v in out.items()})
