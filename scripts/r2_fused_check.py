import sys, torch
sys.path.insert(0, ".")
import bench
from aidet_b200 import _lib as L, synth
from aidet_b200.ops import functional as F
dev = torch.device("cuda", 0)
cb, cs, cg, ng = bench.nms_inputs(dense=False, images=1)
cbd, csd, cgd = cb.to(dev), cs.to(dev), cg.to(dev)
def launches(tag, **kw):
    l0 = L.launch_count(); k = F.nms_batched(cbd, csd, cgd, 0.5, n_groups=ng, **kw); torch.cuda.synchronize()
    print(tag, "launches", L.launch_count() - l0, "kept", (k[0] if isinstance(k, tuple) else k).shape)
launches("default")
L.prof_enable(True); launches("prof on"); L.prof_enable(False)
mine = torch.ones_like(cg, dtype=torch.bool)
cbd, csd, cgd = cb[mine].to(dev), cs[mine].to(dev), cg[mine].to(dev)
launches("after bool-index")
print("ptr align", cbd.data_ptr() % 16, csd.data_ptr() % 16, cgd.data_ptr() % 16, cgd.dtype)
ffma = L.ffma_peak_tflops(0, 1024); launches("after ffma")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_(); launches("after flush alloc")
big = torch.empty((100000, 100000), dtype=torch.float32, device=dev); launches("with 40 GB allocated")
