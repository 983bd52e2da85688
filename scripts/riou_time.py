"""Time the rotated-IoU matrix kernel (dense roofline set and DOTA-shaped sparse set, 32768^2) and check a 512x512
block against the float64 oracle.  Used with AIDET_B200_LIB=<variant .so> for kernel-tuning builds."""
import os
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.ops import functional as F
from oracle import oracle as O
dev = torch.device("cuda", 0)
n = 32768
res = []
for dense, side in ((True, 16384), (False, 1024), (False, 16384)):
    a, _ = synth.dota_boxes(n, side=side, seed=0, dense=dense)
    b, _ = synth.dota_boxes(n, side=side, seed=1, dense=dense)
    ad, bd = a.to(dev), b.to(dev)
    out = torch.empty((n, n), device=dev)
    for _ in range(2):
        F.riou_matrix(ad, bd, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        F.riou_matrix(ad, bd, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    err = abs(out[:512, :512].cpu().double().numpy() - O.riou_matrix(a[:512].numpy(), b[:512].numpy())).max()
    res.append("%s %.1f Gpairs/s err %.1e" % ("dense" if dense else "sparse(side %d)" % side, n * n / ms / 1e6, err))
    del out
print(os.environ.get("AIDET_B200_LIB", "default"), "|", " | ".join(res))
