"""One dense 16384-box group (sort + mask + scan path) and the 8-tile batch: device time of the whole call."""
import os, sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(tag, b, s, g, ng):
    b, s = b.to(dev), s.to(dev); g = None if g is None else g.to(dev)
    for _ in range(3): k = F.nms_batched(b, s, g, 0.5, n_groups=ng)
    L.prof_enable(True); L.prof_read(1, reset=True)
    for _ in range(10):
        flush.zero_(); F.nms_batched(b, s, g, 0.5, n_groups=ng)
    torch.cuda.synchronize()
    ms, n = L.prof_read(1, reset=True); L.prof_enable(False)
    print("%s %s: n=%d kept=%d device %.4f ms" % (os.environ.get("AIDET_B200_LIB", "default")[-24:], tag, b.shape[0], k.shape[0], ms / n))
bb, bs = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
run("one_group_dense", bb, bs, None, 1)
run("c2x8_dense", *bench.nms_inputs(dense=True, images=8))
run("c2x8", *bench.nms_inputs(dense=False, images=8))
