"""One launch of every hot kernel at its bench size, for `ncu --set full` captures:
riou 32768^2 dense, batched NMS (C2 and one dense 16384-box group), RoIAlign C3 fwd + gather bwd, the fused max-IoU
assignment (anchor grid of a 1024 tile x 128 truths) and the rotated-IoU-loss gradient kernel (65536 pairs)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
n = 32768
a, _ = synth.dota_boxes(n, side=16384, seed=0, dense=True)
b, _ = synth.dota_boxes(n, side=16384, seed=1, dense=True)
a, b = a.to(dev), b.to(dev)
out = torch.empty((n, n), device=dev)
F.riou_matrix(a, b, out=out)
del out
cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
F.nms_batched(cb.to(dev), cs.to(dev), cg.to(dev), 0.5, n_groups=ng)
ob, osc = synth.dota_boxes(16384, side=1024, seed=7, dense=True)
F.nms_batched(ob.to(dev), osc.to(dev), None, 0.5, n_groups=1)
feats = [f.to(dev) for f in synth.fpn_features()]
rois, lvl = synth.rotated_rois()
rois, lvl = rois.to(dev), lvl.to(dev)
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
o = F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl)
go = torch.randn_like(o)
grads = [torch.empty_like(f) for f in feats]
F.rroi_align_backward_gather(go, grads, rois, scales, 2, 2, lvl)
from aidet_b200.core import MaxIoUAssigner
_, gt, _, lab = synth.assign_case(1024, 128, seed=21)
MaxIoUAssigner(0.7, 0.3, 0.3, True).assign(synth.anchor_grid().to(dev), gt.to(dev), None, lab.to(dev))
pr, tg = synth.regression_pairs(65536, seed=22)
F.riou_aligned_grad(pr.to(dev), tg.to(dev))
torch.cuda.synchronize()
