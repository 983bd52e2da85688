"""On-GPU probe of batched NMS (used under ncu for launch lists): python scripts/nms_probe.py [images] [dense] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
images = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dense = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cb, cs, cg, ng = bench.nms_inputs(dense=dense, images=images)
cb, cs, cg = cb.to(dev), cs.to(dev), cg.to(dev)
for _ in range(reps):
    k = F.nms_batched(cb, cs, cg, 0.5, n_groups=ng)
torch.cuda.synchronize()
print(cb.shape[0], k.shape[0])
