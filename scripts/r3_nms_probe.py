"""One call of every NMS configuration of the final kernels, for ncu captures: C2 and C1 (fused kernel), 8 tiles batched
(bucket rank + warp-level mask units + scan), one dense 16384-box group (bucket rank + ticketed mask tiles + scan)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
a1, s1 = synth.dota_boxes(2000, side=1024, seed=0)
xb, xs, xg, xng = bench.nms_inputs(dense=True, images=8)
ob, osc = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
for _ in range(2):
    F.nms_batched(cb.to(dev), cs.to(dev), cg.to(dev), 0.5, n_groups=ng)
    F.nms_batched(a1.to(dev), s1.to(dev), None, 0.1, n_groups=1)
    F.nms_batched(xb.to(dev), xs.to(dev), xg.to(dev), 0.5, n_groups=xng)
    F.nms_batched(ob.to(dev), osc.to(dev), None, 0.5, n_groups=1)
torch.cuda.synchronize()
