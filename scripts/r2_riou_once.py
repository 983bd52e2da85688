"""One dense rotated-IoU matrix launch (16384^2) for an ncu capture of riou_matrix_kernel<RectKind>."""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
n = 16384
a, _ = synth.dota_boxes(n, side=16384, seed=0, dense=True)
b, _ = synth.dota_boxes(n, side=16384, seed=1, dense=True)
ad, bd = a.to(dev), b.to(dev)
out = torch.empty((n, n), device=dev)
for _ in range(3):
    F.riou_matrix(ad, bd, out=out)
torch.cuda.synchronize()
