"""Time batched NMS configs on the GPU: whole call (CUDA events, median) and the mask kernel alone (library events)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(name, cb, cs, cg, ng, iters=20):
    cb, cs = cb.to(dev), cs.to(dev)
    cg = cg.to(dev) if cg is not None else None
    for _ in range(3):
        k = F.nms_batched(cb, cs, cg, 0.5, n_groups=ng)
    ts = []
    L.prof_enable(True); L.prof_read(1, reset=True)
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); k = F.nms_batched(cb, cs, cg, 0.5, n_groups=ng); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms_k, nl = L.prof_read(1, reset=True); L.prof_enable(False)
    ts.sort()
    pairs = bench.group_pairs(cg.cpu() if cg is not None else torch.zeros(cb.shape[0], dtype=torch.int32), ng)
    kms = ms_k / max(nl, 1)
    print("%-16s n=%6d kept=%6d call %.4f ms  mask %.4f ms  frac %.3f" % (name, cb.shape[0], k.shape[0], ts[len(ts) // 2], kms,
          pairs * 256 / (kms * 1e-3) / 74.45e12))
run("c2", *bench.nms_inputs(dense=False, images=1))
run("c2_dense", *bench.nms_inputs(dense=True, images=1))
run("c2x8_dense", *bench.nms_inputs(dense=True, images=8))
ob, osc = synth.dota_boxes(16384, side=1024, seed=7, dense=True)
run("one_group_dense", ob, osc, None, 1, iters=10)
