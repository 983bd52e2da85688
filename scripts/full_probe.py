"""One launch of every hot kernel at its bench size, for `ncu --set full` captures (round 2 kernel set):
riou 32768^2 dense -- theta-OBB, rectangles given as 8 points (parallelogram path), free quads (general fan path) and the
TMA tensor-store variant with two destinations; fused batched NMS (C2, C1) and the three-kernel path on one dense
16384-box group; RoIAlign C3 forward (chunk-pair tap-list kernel) + gather backward; fused max-IoU assignment."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
n = 32768
a, _ = synth.dota_boxes(n, side=16384, seed=4, dense=True)
b, _ = synth.dota_boxes(n, side=16384, seed=5, dense=True)
out = torch.empty((n, n), device=dev)
F.riou_matrix(a.to(dev), b.to(dev), out=out)
a8, b8 = synth.thetaobb2pointobb(a).float().to(dev), synth.thetaobb2pointobb(b).float().to(dev)
F.riou_matrix(a8, b8, out=out)
f8a, f8b = synth.free_quads(a, 0.03, seed=7)[0].float().to(dev), synth.free_quads(b, 0.03, seed=8)[0].float().to(dev)
F.riou_matrix(f8a, f8b, out=out)
out2 = torch.empty((n, n), device=dev)
F.riou_matrix_multi(a.to(dev), b.to(dev), [out.data_ptr(), out2.data_ptr()], n)
del out, out2
cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
F.nms_batched(cb.to(dev), cs.to(dev), cg.to(dev), 0.5, n_groups=ng)
a1, s1 = synth.dota_boxes(2000, side=1024, seed=0)
F.nms_batched(a1.to(dev), s1.to(dev), None, 0.1, n_groups=1)
ob, osc = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
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
torch.cuda.synchronize()
