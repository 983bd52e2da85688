"""Device time per call of the batched (large-path) NMS configs: 8 tiles batched, DOTA-shaped and dense."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L
from aidet_b200.ops import functional as F

dev = torch.device("cuda", 0)
for tag, dense in (("c2x8", False), ("c2x8_dense", True)):
    b, s, g, ng = bench.nms_inputs(dense=dense, images=8)
    b, s, g = b.to(dev), s.to(dev), g.to(dev)
    ws = torch.zeros(L.lib().aidet_nms_workspace_bytes(b.shape[0], ng, b.shape[1]) + 128, dtype=torch.uint8, device=dev)
    for _ in range(5):
        keep, nk = F.nms_batched(b, s, g, 0.5, n_groups=ng, sync=False, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        F.nms_batched(b, s, g, 0.5, n_groups=ng, sync=False, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    print("%-11s n=%6d groups=%4d kept=%6d  %.1f us per call" % (tag, b.shape[0], ng, int(nk), e0.elapsed_time(e1) * 20))
