"""On-GPU probe of the fused max-IoU assignment (used under ncu for launch lists):
python scripts/assign_probe.py [n_boxes] [n_gts] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth
from aidet_b200.core import MaxIoUAssigner
dev = torch.device("cuda", 0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 261888
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bx, gt, _, lab = synth.assign_case(nb, ng, seed=21)
bx, gt, lab = bx.to(dev), gt.to(dev), lab.to(dev)
a = MaxIoUAssigner(0.7, 0.3, 0.3, True)
for _ in range(reps):
    r = a.assign(bx, gt, None, lab)
torch.cuda.synchronize()
print(nb, ng, int((r.gt_inds > 0).sum()), int((r.gt_inds == 0).sum()), int((r.gt_inds < 0).sum()))
