"""TMA-store variant of the overlap-matrix kernel with ONE local destination vs the plain kernel (compute speed of the variant)."""
import sys, torch
sys.path.insert(0, ".")
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
n = 32768
a, _ = synth.dota_boxes(n, side=16384, seed=4, dense=True); b, _ = synth.dota_boxes(n, side=16384, seed=5, dense=True)
a, b = a.to(dev), b.to(dev)
out = torch.empty((n, n), device=dev); out2 = torch.empty((n, n), device=dev)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
p = t(lambda: F.riou_matrix(a, b, out=out))
q = t(lambda: F.riou_matrix_multi(a, b, [out2.data_ptr()], n))
q2 = t(lambda: F.riou_matrix_multi(a, b, [out2.data_ptr(), out.data_ptr()], n))
print("plain %.3f ms (%.1f Gpairs/s) | tma 1 dest %.3f ms (%.1f) | tma 2 local dests %.3f ms (%.1f) | equal %s"
      % (p, n * n / p / 1e6, q, n * n / q / 1e6, q2, n * n / q2 / 1e6, bool(torch.equal(out, out2))))
