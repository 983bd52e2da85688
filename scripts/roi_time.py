"""Time the C3 RoIAlign forward / gather backward on the GPU (CUDA events, L2 flushed between iterations)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, iters=20):
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
ws = [None]
def bwd():
    ws[0] = F.rroi_align_backward_gather(go, grads, rois, scales, 2, 2, lvl, workspace=ws[0])
for _ in range(3):
    F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl, out=out); bwd()
print("fwd ms", timed(lambda: F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl, out=out)), "bwd ms", timed(bwd))
