"""Upper bound of what RoI ordering buys the C3 RoIAlign forward / backward: the same RoIs in their given order
(image-major, random positions) and sorted by (level, image, 64-px cell row, x)."""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
feats = [f.to(dev) for f in synth.fpn_features()]
rois, lvl = synth.rotated_rois()
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
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


for name, order in (("given", torch.arange(rois.shape[0])),
                    ("sorted", torch.argsort(lvl.double() * 1e9 + rois[:, 0].double() * 1e7 + (rois[:, 2] // 64).double() * 1e4 + rois[:, 1].double()))):
    r, l = rois[order].contiguous().to(dev), lvl[order].contiguous().to(dev)
    out = F.rroi_align_forward(feats, r, scales, (7, 7), 2, 2, l)
    go = torch.randn_like(out)
    grads = [torch.empty_like(f) for f in feats]
    ws = [None]

    def bwd():
        ws[0] = F.rroi_align_backward_gather(go, grads, r, scales, 2, 2, l, workspace=ws[0])
    for _ in range(3):
        F.rroi_align_forward(feats, r, scales, (7, 7), 2, 2, l, out=out); bwd()
    print(name, "fwd ms", timed(lambda: F.rroi_align_forward(feats, r, scales, (7, 7), 2, 2, l, out=out)), "bwd ms", timed(bwd))
