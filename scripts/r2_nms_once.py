"""One dense 16384-box group through the large NMS path, for an ncu capture of nms_mask_kernel / nms_scan_kernel."""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
bb, bs = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
bb, bs = bb.to(dev), bs.to(dev)
for _ in range(3):
    F.nms_batched(bb, bs, None, 0.5)
torch.cuda.synchronize()
