"""Phase times of the fused NMS kernel (clock stamps of CTA 0 left at the end of the workspace) for C1 / C2 / dense."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F

dev = torch.device("cuda", 0)
NAMES = ["rank", "sync1", "mask", "sync2", "scan(cta0)", "tail(last cta)"]


def run(tag, boxes, scores, groups, thr, ng):
    boxes, scores = boxes.to(dev), scores.to(dev)
    groups = None if groups is None else groups.to(dev)
    n = boxes.shape[0]
    nbytes = L.lib().aidet_nms_workspace_bytes(n, ng, boxes.shape[1])
    ws = torch.zeros(nbytes + 128, dtype=torch.uint8, device=dev)
    L.prof_enable(2)
    for _ in range(5):
        keep, nk = F.nms_batched(boxes, scores, groups, thr, n_groups=ng, sync=False, workspace=ws)
    torch.cuda.synchronize()
    L.prof_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        F.nms_batched(boxes, scores, groups, thr, n_groups=ng, sync=False, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    L.prof_enable(2)
    F.nms_batched(boxes, scores, groups, thr, n_groups=ng, sync=False, workspace=ws)
    base = (ws.data_ptr() + 127) // 128 * 128 - ws.data_ptr()
    tail = ws[base + nbytes - 256: base + nbytes].view(torch.int64)
    tail[16:24] = 0                                                   # max-over-CTAs stamps of ONE call
    F.nms_batched(boxes, scores, groups, thr, n_groups=ng, sync=False, workspace=ws)
    torch.cuda.synchronize()
    st = tail.cpu().tolist()
    L.prof_enable(0)
    t0 = (~st[16]) & ((1 << 64) - 1) if st[16] < 0 else ~st[16]
    gt = [st[16 + k] for k in range(1, 6)]
    print("           latest CTA at each boundary, us after the first CTA started: " + "  ".join("%.1f" % ((g - t0) / 1e3) for g in gt))
    d = [(st[i + 1] - st[i]) / 1965.0 for i in range(5)] + [(st[6] - st[5]) / 1965.0 if st[7] == 0 else float("nan")]
    print("%-10s n=%5d groups=%3d kept=%5d  call %.1f us | " % (tag, n, ng, int(nk), e0.elapsed_time(e1) * 20)
          + "  ".join("%s %.1f" % (a, b) for a, b in zip(NAMES, d)) + "  (us at 1965 MHz)")
    print("           key build of CTA 0: %.1f us" % ((st[13] - st[0]) / 1965.0))
    print("           mask phase of CTA 0: bounds %.1f  unit prefix %.1f  first unit of warp 0 %.1f  warp 0 done %.1f  all warps %.1f us"
          % tuple((st[b] - st[a]) / 1965.0 for a, b in ((2, 14), (14, 15), (15, 24), (24, 25), (25, 3))))
    print("           first unit of warp 0: index math %.2f  loads %.2f  rows %.2f  store %.2f us"
          % tuple((st[b] - st[a]) / 1965.0 for a, b in ((15, 26), (26, 27), (27, 28), (28, 24))))
    if st[12] > 0:
        print("           scan of CTA 0's group, cycles per block over %d blocks: helper warp 0 waits for the keep word %.0f, for the block's copy %.0f (of which until the producer has issued it %.0f), works %.0f; the chain waits for the helpers %.0f"
              % (st[12], st[8] / st[12], st[9] / st[12], st[29] / st[12], st[10] / st[12], st[11] / st[12]))
        print("           longest wait for an issue: %d cycles at block %d; %d blocks waited > 200 cycles" % (st[30], st[31] // 1000, st[31] % 1000))
        print("           producer: %d cycles in all, %d waiting for helpers (%d sleeps), %d for landings, D = %d" % (st[27], st[24], st[26], st[25], st[28]))
    if False:
        print("           chain of group 0, cycles per block over %d blocks: wait-helpers %.0f  wait-data %.0f  bits %.0f  publish %.0f"
              % (st[12], st[8] / st[12], st[9] / st[12], st[10] / st[12], st[11] / st[12]))


a1, s1 = synth.dota_boxes(2000, side=1024, seed=0)
run("C1", a1, s1, None, 0.1, 1)
cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
run("C2", cb, cs, cg, 0.5, ng)
cb, cs, cg, ng = bench.nms_inputs(dense=True, images=1)
run("C2 dense", cb, cs, cg, 0.5, ng)
ob, osc = synth.dota_boxes(8192, side=16384, seed=11, dense=True)
run("8k dense", ob, osc, None, 0.5, 1)
