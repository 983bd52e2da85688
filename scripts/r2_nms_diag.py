import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aidet_b200 import synth, sharded
from aidet_b200.ops import functional as F
from oracle import oracle as O
dev = torch.device("cuda", 0)
bx, sc, lb, ti, org = synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=400, seed=3)
g1 = (ti * 15 + lb).int()
def check(tag, b, s, g, thr, ng):
    keep = F.nms_batched(b.to(dev), s.to(dev), g.to(dev), thr if not torch.is_tensor(thr) else thr.to(dev), n_groups=ng).cpu().numpy()
    ref, near = O.nms(b.numpy(), s.numpy(), thr if not torch.is_tensor(thr) else thr.numpy(), groups=g.numpy(), cmp_ge=False, plus_one=False)
    bad, _ = O.nms_verify(b.numpy(), s.numpy(), thr if not torch.is_tensor(thr) else thr.numpy(), keep, groups=g.numpy(), cmp_ge=False, plus_one=False)
    cnt = np.bincount(g.numpy(), minlength=ng)
    print("%-28s n=%d groups=%d max group=%d: kept %d ref %d near %d verify-bad %d equal %s" % (tag, b.shape[0], ng, cnt.max(), keep.size, ref.size, near, bad, np.array_equal(keep, ref)))
    if not np.array_equal(keep, ref):
        d = np.setxor1d(keep, ref)
        print("   differing indices:", d[:10], "groups", g.numpy()[d[:10]], "group sizes", cnt[g.numpy()[d[:10]]])
NG = int(g1.max()) + 1
check("scene stage1 unsorted", bx, sc, g1, 0.5, NG)
o = torch.argsort(g1, stable=True)
check("scene stage1 sorted", bx[o], sc[o], g1[o], 0.5, NG)
check("scene stage1 first 3000", bx[:3000], sc[:3000], g1[:3000], 0.5, NG)
check("scene stage1 groups<60", bx[g1 < 60], sc[g1 < 60], g1[g1 < 60], 0.5, 60)
keep1 = torch.from_numpy(O.nms(bx.numpy(), sc.numpy(), 0.5, groups=g1.numpy(), cmp_ge=False, plus_one=False)[0])
sb = sharded.translate_to_scene(bx[keep1], org[ti[keep1]])
check("scene stage2", sb, sc[keep1], lb[keep1].int(), sharded.merge_thresholds('obb'), 15)
a, s = synth.dota_boxes(4000, side=600, seed=1)
g = torch.randint(0, 135, (4000,), generator=torch.Generator().manual_seed(0)).int()
check("random groups", a, s, g, 0.3, 135)
check("random groups sorted", a[torch.argsort(g, stable=True)], s[torch.argsort(g, stable=True)], g[torch.argsort(g, stable=True)], 0.3, 135)
g2 = (torch.arange(4000) // 1300).int()
check("3 big sorted groups", a, s, g2, 0.3, 4)
check("1 group 2000", a[:2000], s[:2000], torch.zeros(2000, dtype=torch.int32), 0.3, 1)
check("1 group 2100", a[:2100], s[:2100], torch.zeros(2100, dtype=torch.int32), 0.3, 1)
check("1 group 4000", a, s, torch.zeros(4000, dtype=torch.int32), 0.3, 1)
