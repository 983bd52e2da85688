"""On-GPU probe of the RoIAlign C3 kernels (used under ncu for launch lists)."""
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
out = F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl)
go = torch.randn_like(out)
grads = [torch.empty_like(f) for f in feats]
ws = None
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl, out=out)
    ws = F.rroi_align_backward_gather(go, grads, rois, scales, 2, 2, lvl, workspace=ws)
torch.cuda.synchronize()
