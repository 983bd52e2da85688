"""GPU parity: NMS keep indices vs the oracle.

Bar (BASELINE.json north_star): keep indices bit-exact, excluding pairs whose IoU lies within
1e-6 of the threshold; those pairs are counted and reported.  When the oracle finds no such
pair among the evaluated ones the comparison is exact; otherwise the keep set must still be
a valid greedy outcome inside the +-1e-6 band (oracle_nms_verify).
"""
import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.core import multiclass_nms, multiclass_nms_with_index, multiclass_thetaobb_nms
from aidet_b200.ops import batched_rnms, nms, pointobb_nms, thetaobb_nms
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _compare(boxes, scores, thr, keep_gpu, groups=None, cmp_ge=False, plus_one=False):
    keep_gpu = keep_gpu.cpu().numpy()
    ref, near = O.nms(boxes.numpy(), scores.numpy(), thr, groups=None if groups is None else groups.numpy(),
                      cmp_ge=cmp_ge, plus_one=plus_one)
    if near == 0:
        assert np.array_equal(keep_gpu, ref), "keep mismatch: %d vs %d kept" % (len(keep_gpu), len(ref))
    else:
        bad, near2 = O.nms_verify(boxes.numpy(), scores.numpy(), thr, keep_gpu,
                                  groups=None if groups is None else groups.numpy(), cmp_ge=cmp_ge,
                                  plus_one=plus_one)
        print("near-threshold pairs reported: %d" % max(near, near2))
        assert bad == 0
    return len(ref), near


@pytest.mark.parametrize("dense", [False, True])
def test_c1_single_class_nms(cuda, dense):
    """Config C1: 2000 theta-OBBs, single class, threshold 0.1 (harbor / tennis-court, dota.py:324)."""
    boxes, scores = synth.dota_boxes(2000, seed=0, dense=dense)
    dets = torch.cat([boxes, scores[:, None]], 1).to(cuda)
    kept, inds = thetaobb_nms(dets, 0.1)
    assert inds.dtype == torch.long and inds.is_cuda and torch.equal(kept, dets[inds])
    assert (inds[1:] > inds[:-1]).all()          # ascending original index (nms_kernel.cu:135-138)
    _compare(boxes, scores, 0.1, inds)


@pytest.mark.parametrize("thr", [0.05, 0.3, 0.5, 0.9])
def test_thresholds(cuda, thr):
    boxes, scores = synth.dota_boxes(1500, side=512, seed=21)
    inds = thetaobb_nms(torch.cat([boxes, scores[:, None]], 1).to(cuda), thr)[1]
    _compare(boxes, scores, thr, inds)


@pytest.mark.parametrize("n,side,thr", [(20000, 300, 0.1), (20000, 500, 0.3), (20000, 16384, 0.5), (9000, 500, 0.1)])
def test_large_single_group(cuda, n, side, thr):
    """One group beyond the fused kernel's 8192 boxes: radix sort + ticketed mask + the scan kernel that jumps over runs of
    dead 32-row blocks (crowded set: most blocks die early), walks live blocks with the prefetch (sparse set: every block
    lives) and reads columns beyond its 16384-column panel (n = 20000)."""
    boxes, scores = synth.dota_boxes(n, side=side, seed=27)
    inds = thetaobb_nms(torch.cat([boxes, scores[:, None]], 1).to(cuda), thr)[1]
    kept, _ = _compare(boxes, scores, thr, inds)
    print("n=%d side=%d: %d kept" % (n, side, kept))


@pytest.mark.parametrize("fmt,cmp_ge", [(5, False), (8, False), (5, True), (4, True)])
def test_many_small_groups_beyond_the_fused_kernel(cuda, fmt, cmp_ge):
    """n > 8192 in groups of a few hundred boxes (batched tiles, a scene's tile x class groups): radix sort + the warp-level
    mask units (nms_mask_units_kernel) + the scan kernel; ragged sizes, one empty group, per-group thresholds, ids outside
    [0, n_groups) (never scanned, never kept) and keep_fill."""
    from aidet_b200.ops import functional as F
    n, ng = 12000, 40
    boxes, scores = synth.dota_boxes(n, side=900, seed=61)
    g = torch.Generator().manual_seed(5)
    groups = torch.randint(0, ng - 1, (n,), generator=g).int()           # group ng-1 stays empty
    groups[torch.randperm(n, generator=g)[:300]] = ng + 3                # not this call's share
    thr = torch.linspace(0.2, 0.6, ng)
    if fmt == 8:
        boxes = synth.thetaobb2pointobb(boxes).float()
    elif fmt == 4:
        p = synth.thetaobb2pointobb(boxes).view(n, 4, 2)
        boxes = torch.cat([p.min(1).values, p.max(1).values], 1).float()
    keep, nk = F.nms_batched(boxes.to(cuda), scores.to(cuda), groups.to(cuda), thr.to(cuda), n_groups=ng, cmp_ge=cmp_ge,
                             plus_one=(fmt == 4), sync=False, keep_fill=n)
    k = int(nk)
    assert (keep[k:] == n).all()
    inds = keep[:k].cpu()
    assert (inds[1:] > inds[:-1]).all() and (groups[inds] < ng).all()
    ok = groups < ng
    sel = ok.nonzero().flatten()
    pos = torch.full((n,), -1, dtype=torch.long); pos[sel] = torch.arange(sel.numel())
    ref_total = 0
    for c in range(ng):                                                   # the oracle class by class (per-group thresholds)
        m = (groups == c).nonzero().flatten()
        if m.numel() == 0:
            continue
        got_c = inds[groups[inds] == c]
        local = torch.searchsorted(m, got_c)
        assert torch.equal(m[local], got_c)
        kept, _ = _compare(boxes[m], scores[m], float(thr[c]), local.to(cuda), cmp_ge=cmp_ge, plus_one=(fmt == 4))
        ref_total += kept
    print("fmt %d ge %d: %d kept of %d" % (fmt, cmp_ge, k, int(ok.sum())))


def test_pointobb_nms(cuda):
    boxes, scores = synth.dota_boxes(1200, side=512, seed=22)
    p8 = synth.thetaobb2pointobb(boxes)
    inds = pointobb_nms(torch.cat([p8, scores[:, None]], 1).to(cuda), 0.3)[1]
    _compare(p8, scores, 0.3, inds)


def test_c2_batched_15_classes(cuda):
    """Config C2: 2000 proposals x 15 classes, thr 0.5, score_thr 0.05, max_per_img 1000."""
    mb, ms = synth.multiclass_dets(2000, 15, seed=2)
    dets, labels = multiclass_thetaobb_nms(mb.to(cuda), ms.to(cuda), 0.05, 0.5, 1000)
    assert dets.shape[1] == 6 and labels.dtype == torch.long and dets.shape[0] <= 1000
    # oracle: per-class greedy NMS on the same filtered candidates, class-major, top-1000 by score
    n, C = ms.shape[0], ms.shape[1] - 1
    boxes = mb.view(n, C + 1, 5)[:, 1:]
    valid = (ms[:, 1:] > 0.05).t()
    lab, rows = valid.nonzero(as_tuple=True)
    cand_b, cand_s = boxes[rows, lab], ms[:, 1:][rows, lab]
    keep_gpu = batched_rnms(cand_b.to(cuda), cand_s.to(cuda), lab.to(cuda), 0.5, n_groups=C)
    nref, near = _compare(cand_b, cand_s, 0.5, keep_gpu, groups=lab.int())
    ref_keep, _ = O.nms(cand_b.numpy(), cand_s.numpy(), 0.5, groups=lab.int().numpy())
    if near == 0:
        ref_d = torch.cat([cand_b[ref_keep], cand_s[ref_keep, None]], 1)
        ref_l = lab[ref_keep]
        if ref_d.shape[0] > 1000:
            order = ref_d[:, -1].sort(descending=True)[1][:1000]
            ref_d, ref_l = ref_d[order], ref_l[order]
        assert torch.equal(dets.cpu(), ref_d) and torch.equal(labels.cpu(), ref_l)


def test_per_group_thresholds_and_empty_groups(cuda):
    boxes, scores = synth.dota_boxes(3000, side=700, seed=23)
    g = torch.Generator().manual_seed(3)
    groups = torch.randint(0, 15, (3000,), generator=g).int()
    groups[groups == 4] = 5                       # group 4 is empty
    thr = torch.tensor([0.1, 0.05, 0.15, 0.5, 0.35, 0.2, 0.2, 0.45, 0.2, 0.1, 0.1, 0.15, 0.05, 0.4, 0.2])
    keep = batched_rnms(boxes.to(cuda), scores.to(cuda), groups.to(cuda), thr.to(cuda), n_groups=15)
    _compare(boxes, scores, thr.numpy(), keep, groups=groups)


def test_ragged_group_sizes(cuda):
    """groups of 1, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000 boxes."""
    sizes = [1, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000]
    boxes, scores = synth.dota_boxes(sum(sizes), side=256, seed=24)
    groups = torch.cat([torch.full((s,), i, dtype=torch.int32) for i, s in enumerate(sizes)])
    perm = torch.randperm(sum(sizes), generator=torch.Generator().manual_seed(1))
    boxes, scores, groups = boxes[perm], scores[perm], groups[perm]
    keep = batched_rnms(boxes.to(cuda), scores.to(cuda), groups.to(cuda), 0.3)
    _compare(boxes, scores, 0.3, keep, groups=groups)


def test_duplicates_and_ties(cuda):
    """identical boxes and tied scores: order must be (score desc, original index asc)."""
    boxes, scores = synth.dota_boxes(500, side=256, seed=25)
    boxes = torch.cat([boxes, boxes[:200]])
    scores = torch.cat([scores, scores[:200]])
    scores[::7] = 0.5
    keep = batched_rnms(boxes.to(cuda), scores.to(cuda), None, 0.5)
    _compare(boxes, scores, 0.5, keep)


def test_hbb_nms_reference_vectors(cuda):
    """tests/test_nms.py:16-41 and nms_wrapper.py:25-34 known answers (+1 convention)."""
    base = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9],
                     [35.3, 11.5, 39.9, 14.5, 0.4], [35.2, 11.7, 39.7, 15.7, 0.3]])
    for dt in (np.float32, np.float64):
        sup, inds = nms(base.astype(dt), 0.7)
        assert isinstance(sup, np.ndarray) and sup.dtype == dt and inds.dtype == np.int64
        assert len(inds) == len(sup) == 3 and list(inds) == [0, 2, 3]
    for tt in (torch.FloatTensor, torch.DoubleTensor):
        t = tt(base)
        sup, inds = nms(t, 0.7)
        assert sup.dtype == t.dtype and inds.dtype == torch.long and not inds.is_cuda and len(inds) == 3
    sup, inds = nms(torch.tensor(base, dtype=torch.float32, device=cuda), 0.7)
    assert inds.is_cuda and inds.tolist() == [0, 2, 3]
    seven = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9], [49.2, 31.8, 51.0, 35.4, 0.5],
                      [35.1, 11.5, 39.1, 15.7, 0.5], [35.6, 11.8, 39.3, 14.2, 0.5], [35.3, 11.5, 39.9, 14.5, 0.4],
                      [35.2, 11.7, 39.7, 15.7, 0.3]], dtype=np.float32)
    sup, inds = nms(seven, 0.7)
    assert len(inds) == len(sup) == 3
    sup, inds = nms(seven, 0.7, device_id=0)
    assert isinstance(inds, np.ndarray) and len(inds) == 3
    empty = torch.zeros((0, 5), device=cuda)
    sup, inds = nms(empty, 0.5)
    assert sup.shape == (0, 5) and inds.shape == (0,) and inds.dtype == torch.long
    with pytest.raises(TypeError):
        nms([[1, 2, 3, 4, 0.5]], 0.5)


def test_hbb_nms_random_vs_oracle(cuda):
    boxes5, scores = synth.dota_boxes(2000, seed=26)
    x1 = boxes5[:, 0] - boxes5[:, 2] / 2
    y1 = boxes5[:, 1] - boxes5[:, 3] / 2
    hbb = torch.stack([x1, y1, x1 + boxes5[:, 2], y1 + boxes5[:, 3]], 1)
    dets = torch.cat([hbb, scores[:, None]], 1)
    inds = nms(dets.to(cuda), 0.5)[1]                      # CUDA semantics: >
    _compare(hbb, scores, 0.5, inds, cmp_ge=False, plus_one=True)
    inds_cpu_sem = nms(dets, 0.5)[1]                       # CPU-tensor entry: >=
    _compare(hbb, scores, 0.5, inds_cpu_sem, cmp_ge=True, plus_one=True)


def test_multiclass_hbb_drivers(cuda):
    g = torch.Generator().manual_seed(5)
    n, C = 600, 4
    boxes5, _ = synth.dota_boxes(n, side=300, seed=27)
    x1 = boxes5[:, 0] - boxes5[:, 2] / 2
    y1 = boxes5[:, 1] - boxes5[:, 3] / 2
    hbb = torch.stack([x1, y1, x1 + boxes5[:, 2], y1 + boxes5[:, 3]], 1)
    mb = (hbb[:, None, :] + torch.randn(n, C + 1, 4, generator=g)).reshape(n, -1)
    ms = torch.softmax(2 * torch.randn(n, C + 1, generator=g), 1)
    dets, labels = multiclass_nms(mb.to(cuda), ms.to(cuda), 0.05, dict(type='nms', iou_thr=0.5), 100)
    d2, l2, cls_inds, keep_inds = multiclass_nms_with_index(mb.to(cuda), ms.to(cuda), 0.05,
                                                            dict(type='nms', iou_thr=0.5), 100)
    assert torch.equal(dets, d2) and torch.equal(labels, l2) and dets.shape[0] <= 100
    assert len(cls_inds) == C
    # per-class check against the oracle
    it = iter(keep_inds)
    for c in range(C):
        m = ms[:, c + 1] > 0.05
        if not m.any():
            continue
        k = next(it).cpu().numpy()
        ref, near = O.nms(mb.view(n, C + 1, 4)[m, c + 1].numpy(), ms[m, c + 1].numpy(), 0.5, plus_one=True)
        if near == 0:
            assert np.array_equal(k, ref)
