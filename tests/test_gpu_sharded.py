"""GPU: aidet_b200.sharded through the CUDA library (single process = world size 1; the multi-rank plumbing is
covered on CPU by tests/test_sharded_gloo.py and on 2-8 GPUs by bench.py under torchrun)."""
import numpy as np
import pytest
import torch

from aidet_b200 import sharded, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_sharded_overlaps_single_rank(cuda):
    a, _ = synth.dota_boxes(300, side=300, seed=1)
    b, _ = synth.dota_boxes(257, side=300, seed=2)
    got = sharded.sharded_rbbox_overlaps(a.to(cuda), b.to(cuda)).cpu().numpy()
    assert np.abs(got - O.riou_matrix(a.numpy(), b.numpy())).max() <= 1e-5


def test_c5_scene_merge_vs_oracle(cuda):
    """Config C5 in miniature: 9 tiles of a 1500 px scene, per-tile NMS @0.5 then class-wise merge (dota.py:324)."""
    bx, sc, lb, ti, org = synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=400, seed=3)
    mb, ms, ml = sharded.scene_merge_nms(bx.to(cuda), sc.to(cuda), lb.to(cuda), ti.to(cuda), org.to(cuda))
    g1 = (ti * 15 + lb).int().numpy()
    keep1, near1 = O.nms(bx.numpy(), sc.numpy(), 0.5, groups=g1, cmp_ge=False, plus_one=False)
    keep1 = torch.from_numpy(keep1)
    sb = sharded.translate_to_scene(bx[keep1], org[ti[keep1]])
    thr = sharded.merge_thresholds('obb').numpy()
    keep2, near2 = O.nms(sb.numpy(), sc[keep1].numpy(), thr, groups=lb[keep1].int().numpy(), cmp_ge=False, plus_one=False)
    keep2 = torch.from_numpy(keep2)
    order = torch.argsort(lb[keep1][keep2], stable=True)
    print("near-threshold pairs reported: %d + %d" % (near1, near2))
    if near1 == 0 and near2 == 0:
        assert torch.equal(mb.cpu(), sb[keep2][order])
        assert torch.equal(ms.cpu(), sc[keep1][keep2][order])
        assert torch.equal(ml.cpu(), lb[keep1][keep2][order])
    else:
        assert abs(mb.shape[0] - len(keep2)) <= near1 + near2
    assert mb.shape[0] < len(keep1) < bx.shape[0]


def test_c5_pointobb_scene_merge(cuda):
    bx, sc, lb, ti, org = synth.scene_dets(scene=1200, tile=512, overlap=100, dets_per_tile=200, seed=4)
    p8 = synth.thetaobb2pointobb(bx)
    m5 = sharded.scene_merge_nms(bx.to(cuda), sc.to(cuda), lb.to(cuda), ti.to(cuda), org.to(cuda))
    m8 = sharded.scene_merge_nms(p8.to(cuda), sc.to(cuda), lb.to(cuda), ti.to(cuda), org.to(cuda))
    assert m8[0].shape[1] == 8
    # same detections survive unless a pair sits within f32 rounding of a threshold (8-point inputs are rounded)
    assert abs(m5[0].shape[0] - m8[0].shape[0]) <= 2


def test_scene_merge_native_equals_the_composed_calls(cuda):
    """The library's three scene entry points against the same stages composed from nms_batched in torch (the form
    the multi-tensor-threshold path still takes): identical outputs, bit for bit."""
    bx, sc, lb, ti, org = [t.to(cuda) for t in synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=300, seed=8)]
    native = sharded.scene_merge_nms(bx, sc, lb, ti, org)
    composed = sharded.scene_merge_nms(bx, sc, lb, ti, org, tile_iou_thr=torch.tensor([0.5], device=cuda))
    for a, b in zip(native, composed):
        assert torch.equal(a, b)


def test_scene_merge_empty_and_out_of_range(cuda):
    bx, sc, lb, ti, org = [t.to(cuda) for t in synth.scene_dets(scene=1200, tile=512, overlap=100, dets_per_tile=150, seed=9)]
    eb, es, el = sharded.scene_merge_nms(bx[:0], sc[:0], lb[:0], ti[:0], org)
    assert eb.shape == (0, 5) and es.numel() == 0 and el.numel() == 0
    # detections whose label or tile id is out of range are never kept and do not disturb the others
    bad = torch.zeros_like(lb, dtype=torch.bool)
    bad[::7] = True
    lb2 = torch.where(bad, torch.full_like(lb, 99), lb)
    ti2 = ti.clone()
    ti2[3::11] = -1
    bad |= ti2 < 0
    got = sharded.scene_merge_nms(bx, sc, lb2, ti2, org)
    ok = ~bad
    want = sharded.scene_merge_nms(bx[ok].contiguous(), sc[ok].contiguous(), lb[ok].contiguous(), ti[ok].contiguous(), org)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_scene_merge_many_tiles_few_boxes(cuda):
    """<= 8192 boxes over more than 1024 tile x class groups: stage 1 takes the multi-kernel path, stage 2 (15 groups) the
    fused kernel, whose padded mask is the LARGER workspace layout of the two."""
    bx, sc, lb, ti, org = [t.to(cuda) for t in synth.scene_dets(scene=4400, tile=512, overlap=100, dets_per_tile=30, seed=12)]
    assert org.shape[0] * 15 > 1024 and bx.shape[0] <= 8192
    native = sharded.scene_merge_nms(bx, sc, lb, ti, org)
    composed = sharded.scene_merge_nms(bx, sc, lb, ti, org, tile_iou_thr=torch.tensor([0.5], device=cuda))
    for a, b in zip(native, composed):
        assert torch.equal(a, b)
