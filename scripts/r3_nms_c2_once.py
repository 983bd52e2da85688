"""Config C2 (5822 candidates, 15 classes) through the fused NMS kernel a few times, for an ncu capture of nms_fused_kernel."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from aidet_b200.ops import functional as F

dev = torch.device("cuda", 0)
b, s, g, ng = bench.nms_inputs(dense=False, images=1)
b, s, g = b.to(dev), s.to(dev), g.to(dev)
for _ in range(3):
    F.nms_batched(b, s, g, 0.5, n_groups=ng)
torch.cuda.synchronize()
