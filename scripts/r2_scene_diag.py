import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth, sharded
from aidet_b200.ops import functional as F
from oracle import oracle as O
dev = torch.device("cuda", 0)
bx, sc, lb, ti, org = synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=400, seed=3)
g1 = (ti * 15 + lb).int().numpy()
keep1, near1 = O.nms(bx.numpy(), sc.numpy(), 0.5, groups=g1, cmp_ge=False, plus_one=False)
keep1 = torch.from_numpy(keep1)
sb = sharded.translate_to_scene(bx[keep1], org[ti[keep1]])
thr = sharded.merge_thresholds('obb').numpy()
keep2, near2 = O.nms(sb.numpy(), sc[keep1].numpy(), thr, groups=lb[keep1].int().numpy(), cmp_ge=False, plus_one=False)
keep2 = torch.from_numpy(keep2)
order = torch.argsort(lb[keep1][keep2], stable=True)
want = sb[keep2][order]
print("sorted groups:", bool((np.diff(g1) >= 0).all()), "n", bx.shape[0], "tiles", org.shape[0])
for it in range(6):
    mb, ms, ml = sharded.scene_merge_nms(bx.to(dev), sc.to(dev), lb.to(dev), ti.to(dev), org.to(dev))
    mb = mb.cpu()
    same = mb.shape == want.shape and bool(torch.equal(mb, want))
    print("run %d: shape %s want %s equal %s" % (it, tuple(mb.shape), tuple(want.shape), same), end="")
    if mb.shape == want.shape and not same:
        d = (mb != want).any(1).nonzero().flatten()
        print("  rows differing:", d[:8].tolist(), "max abs diff", float((mb - want).abs().max()))
    else:
        print()
# garbage dependence: poison the cached workspace, then repeat
print("--- after poisoning the cached workspace")
for val in (255, 1, 0):
    for k, w in F._NMS_WS.items():
        w.fill_(val)
    mb, ms, ml = sharded.scene_merge_nms(bx.to(dev), sc.to(dev), lb.to(dev), ti.to(dev), org.to(dev))
    mb = mb.cpu()
    print("fill %3d: shape %s equal %s" % (val, tuple(mb.shape), mb.shape == want.shape and bool(torch.equal(mb, want))))
    g = (ti * 15 + lb).int()
    for val2 in (255,):
        for k, w in F._NMS_WS.items():
            w.fill_(val2)
        k1 = F.nms_batched(bx.to(dev), sc.to(dev), g.to(dev), 0.5, n_groups=240).cpu()
        print("   stage 1 alone after fill %d: equal %s (%d vs %d)" % (val2, bool(torch.equal(k1, keep1)), k1.numel(), keep1.numel()))
        for k, w in F._NMS_WS.items():
            w.fill_(val2)
        k2 = F.nms_batched(sb.to(dev), sc[keep1].to(dev), lb[keep1].int().to(dev), torch.from_numpy(thr).to(dev), n_groups=15).cpu()
        print("   stage 2 alone after fill %d: equal %s (%d vs %d)" % (val2, bool(torch.equal(k2, keep2)), k2.numel(), keep2.numel()))
