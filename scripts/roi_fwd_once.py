"""One C3 RoIAlign forward (after a warm-up) for ncu captures: python scripts/roi_fwd_once.py"""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
feats = [f.to(dev) for f in synth.fpn_features()]
rois, lvl = synth.rotated_rois()
rois, lvl = rois.to(dev), lvl.to(dev)
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
for _ in range(2):
    out = F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl)
torch.cuda.synchronize()
print(float(out.abs().mean()))
