"""8-point (point-OBB) path at the bench sizes: overlap matrix 32768^2 dense, batched NMS C2 and one dense 16384-box
group, kernel times from the library's CUDA-event scopes.  `--once`: one launch of each (for ncu captures)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F

once = "--once" in sys.argv
dev = torch.device("cuda", 0)
res = {}
PEAK = bench.FP32_PEAK_NOMINAL


def timeit(fn, kind, iters=10, warm=3):
    for _ in range(1 if once else warm):
        fn()
    torch.cuda.synchronize()
    if once:
        return 0.0, 0.0
    L.prof_read(kind, reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    k_ms, k_cnt = L.prof_read(kind, reset=True)
    return e0.elapsed_time(e1) / iters, k_ms / max(k_cnt, 1)


L.prof_enable(not once)
n = 32768
for fmt in (5, 8):
    a, _ = synth.dota_boxes(n, side=16384, seed=4, dense=True)
    b, _ = synth.dota_boxes(n, side=16384, seed=5, dense=True)
    if fmt == 8:
        a, b = synth.thetaobb2pointobb(a).float(), synth.thetaobb2pointobb(b).float()
    a, b = a.to(dev), b.to(dev)
    out = torch.empty((n, n), device=dev)
    ms, kms = timeit(lambda: F.riou_matrix(a, b, out=out), L.PROF_RIOU)
    if not once:
        res["riou_dense_fmt%d" % fmt] = {"ms": ms, "kernel_ms": kms, "gpairs_s": n * n / ms / 1e6,
                                         "frac_fp32": n * n * 256.0 / (kms * 1e-3) / 1e12 / PEAK}
    del out
    cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
    if fmt == 8:
        cb = synth.thetaobb2pointobb(cb).float()
    cb, cs, cg = cb.to(dev), cs.to(dev), cg.to(dev)
    ms, kms = timeit(lambda: F.nms_batched(cb, cs, cg, 0.5, n_groups=ng), L.PROF_NMS_MASK, iters=20)
    if not once:
        gp = bench.group_pairs(cg.cpu(), ng)
        res["nms_c2_fmt%d" % fmt] = {"ms": ms, "mask_ms": kms, "mboxes_s": cb.shape[0] / ms / 1e3,
                                     "mask_frac_fp32": gp * 256.0 / (kms * 1e-3) / 1e12 / PEAK,
                                     "call_frac_fp32": gp * 256.0 / (ms * 1e-3) / 1e12 / PEAK}
    ob, osc = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
    if fmt == 8:
        ob = synth.thetaobb2pointobb(ob).float()
    ob, osc = ob.to(dev), osc.to(dev)
    ms, kms = timeit(lambda: F.nms_batched(ob, osc, None, 0.5), L.PROF_NMS_MASK, iters=10)
    if not once:
        gp = 16384 * 16383 / 2.0
        res["nms_one_group_fmt%d" % fmt] = {"ms": ms, "mask_ms": kms, "mboxes_s": 16384 / ms / 1e3,
                                            "mask_frac_fp32": gp * 256.0 / (kms * 1e-3) / 1e12 / PEAK,
                                            "call_frac_fp32": gp * 256.0 / (ms * 1e-3) / 1e12 / PEAK}
if not once:
    print(json.dumps(res, indent=1))
    open("gpurun_out/r2_quad_probe.json", "w").write(json.dumps(res, indent=1))
