"""Where the C5 scene merge spends its time: CUDA kernel list (torch profiler) of one scene_merge_nms call."""
import sys
import torch
sys.path.insert(0, ".")
from aidet_b200 import sharded, synth
dev = torch.device("cuda", 0)
sx, ssc, sl, st, so = [t.to(dev) for t in synth.scene_dets()]
for _ in range(3):
    sharded.scene_merge_nms(sx, ssc, sl, st, so)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    sharded.scene_merge_nms(sx, ssc, sl, st, so)
e1.record(); torch.cuda.synchronize()
print("per call %.3f ms" % (e0.elapsed_time(e1) / 20))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sharded.scene_merge_nms(sx, ssc, sl, st, so)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
tot = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    tot += d
    print("%8.1f us  +%7.1f  %s" % (e.time_range.start - t0, d, e.name[:90]))
print("kernels: %d, sum %.1f us, span %.1f us" % (len(evs), tot, evs[-1].time_range.end - t0))
