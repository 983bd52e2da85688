"""Quick on-GPU probe: FFMA peak, IoU matrix throughput (dense / sparse), NMS and RoIAlign timings."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F

dev = torch.device("cuda", 0)
res = {}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res["ffma_tflops"] = L.ffma_peak_tflops(0, 8192)
for dense in (True, False):
    for n in (2000, 16384, 32768):
        a, _ = synth.dota_boxes(n, side=1024 if n <= 2000 else 16384, seed=0, dense=dense)
        b, _ = synth.dota_boxes(n, side=1024 if n <= 2000 else 16384, seed=1, dense=dense)
        a, b = a.to(dev), b.to(dev)
        out = torch.empty((n, n), device=dev)
        ms = timeit(lambda: F.riou_matrix(a, b, out=out), iters=10)
        res["riou_%s_%d" % ("dense" if dense else "sparse", n)] = {"ms": ms, "gpairs_s": n * n / ms / 1e6}
        del out
for dense in (True, False):
    mb, msc = synth.multiclass_dets(2000, 15, seed=2, dense=dense)
    n, C = msc.shape[0], 15
    boxes = mb.view(n, C + 1, 5)[:, 1:]
    valid = (msc[:, 1:] > 0.05).t()
    lab, rows = valid.nonzero(as_tuple=True)
    cb, cs, lab = boxes[rows, lab].to(dev), msc[:, 1:][rows, lab].to(dev), lab.to(dev)
    ms = timeit(lambda: F.nms_batched(cb, cs, lab, 0.5, n_groups=C), iters=20)
    res["nms_c2_%s" % ("dense" if dense else "sparse")] = {"ms": ms, "boxes": cb.shape[0], "mboxes_s": cb.shape[0] / ms / 1e3}
feats = [f.to(dev) for f in synth.fpn_features()]
rois, lvl = synth.rotated_rois()
rois, lvl = rois.to(dev), lvl.to(dev)
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
ms = timeit(lambda: F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl), iters=20)
res["roi_fwd_ms"] = ms
out = F.rroi_align_forward(feats, rois, scales, (7, 7), 2, 2, lvl)
go = torch.randn_like(out)
grads = [torch.zeros_like(f) for f in feats]
ms = timeit(lambda: F.rroi_align_backward(go, grads, rois, scales, 2, 2, lvl), iters=10)
res["roi_bwd_ms_no_zero"] = ms
print(json.dumps(res, indent=1))
open("gpurun_out/quick_probe.json", "w").write(json.dumps(res, indent=1))
