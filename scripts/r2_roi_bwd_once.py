import sys, torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
feats = [f.to(dev) for f in synth.fpn_features()]
rois, lvl = synth.rotated_rois(); rois, lvl = rois.to(dev), lvl.to(dev)
sc = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
o = F.rroi_align_forward(feats, rois, sc, (7, 7), 2, 2, lvl)
go = torch.randn_like(o); grads = [torch.empty_like(f) for f in feats]
ws = None
for det in (False, True, False, True):
    ws = F.rroi_align_backward_gather(go, grads, rois, sc, 2, 2, lvl, workspace=ws, deterministic=det)
torch.cuda.synchronize()
