"""OBB codecs (aidet_b200.core.rbbox, mirror of mmdet/core/rbbox/transforms.py) and the head post-processing
(get_det_rbboxes, rbbox_head.py:253-296 with the commented rotated NMS of :294-295 enabled)."""
import math

import numpy as np
import pytest
import torch

from aidet_b200 import core


def test_thetaobb2pointobb_reference_fixtures():
    # tests/test_randomflip.py:6-7 fixtures; expected corners = cv2.boxPoints of the reference's own conversion
    # (transforms.py:45-55), values recorded in SURVEY.md 8c(6)
    a = core.thetaobb2pointobb([200, 200, 300, 150, 45 * math.pi / 180])
    b = core.thetaobb2pointobb([700, 800, 300, 200, 135 * math.pi / 180])
    assert np.allclose(a, [40.901, 146.967, 146.967, 40.901, 359.099, 253.033, 253.033, 359.099], atol=2e-3)
    assert np.allclose(b, [735.355, 623.223, 876.777, 764.645, 664.645, 976.777, 523.223, 835.355], atol=2e-3)
    assert core.pointobb2bbox([1, 5, 4, 2, 7, 6, 3, 9]) == [1, 2, 7, 9]                     # transforms.py:57-71


def test_thetaobb2pointobb_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(0)
    t = torch.rand(200, 5, generator=g) * torch.tensor([1000., 1000., 300., 300., math.pi]) - torch.tensor([0, 0, 0, 0, math.pi / 2])
    got = core.thetaobb2pointobb(t.double()).numpy()
    for row, box in zip(got, t.numpy()):
        ref = cv2.boxPoints(((float(box[0]), float(box[1])), (float(box[2]), float(box[3])), float(box[4]) * 180.0 / np.pi))
        assert np.allclose(row, ref.reshape(-1), atol=2e-3)                                   # cv2 works in f32
    assert torch.allclose(core.pointobb2bbox(torch.from_numpy(got))[:, 2], torch.from_numpy(got)[:, 0::2].max(1).values)


def test_delta_codecs_roundtrip_and_formulas():
    g = torch.Generator().manual_seed(1)
    n = 64
    xy = torch.rand(n, 2, generator=g) * 500
    wh = torch.rand(n, 2, generator=g) * 200 + 10
    prop = torch.cat([xy, xy + wh], 1)
    gt = torch.cat([xy + wh / 2 + torch.randn(n, 2, generator=g) * 5, wh * (0.5 + torch.rand(n, 2, generator=g)),
                    (torch.rand(n, 1, generator=g) - 0.5) * math.pi], 1)
    means, stds = [0., 0., 0., 0., 0.], [0.1, 0.1, 0.2, 0.2, 0.1]
    d = core.thetaobb2delta(prop, gt, means, stds)
    # transforms.py:329-346 written out for row 0
    pw, ph = prop[0, 2] - prop[0, 0] + 1, prop[0, 3] - prop[0, 1] + 1
    exp0 = torch.stack([(gt[0, 0] - (prop[0, 0] + prop[0, 2]) / 2) / pw / 0.1, (gt[0, 1] - (prop[0, 1] + prop[0, 3]) / 2) / ph / 0.1,
                        torch.log((gt[0, 2] + 1) / pw) / 0.2, torch.log((gt[0, 3] + 1) / ph) / 0.2, (gt[0, 4] + math.pi / 2) / 0.1])
    assert torch.allclose(d[0], exp0, atol=1e-5)
    back = core.delta2thetaobb(prop, d, means, stds)
    assert torch.allclose(back[:, :2], gt[:, :2], atol=1e-3) and torch.allclose(back[:, 4], gt[:, 4], atol=1e-5)
    assert torch.allclose(back[:, 2:4], gt[:, 2:4] + 1, rtol=1e-4)        # the reference's +1 on w,h is not undone (:339-340,387-388)
    multi = core.delta2thetaobb(prop, d.repeat(1, 3), means, stds)        # class-wise predictions (n, 3*5)
    assert multi.shape == (n, 15) and torch.allclose(multi[:, 5:10], back)
    big = core.delta2thetaobb(prop[:1], torch.tensor([[0., 0., 100., -100., 0.]]))
    assert torch.allclose(big[0, 2], pw * 1000 / 16, rtol=1e-5) and torch.allclose(big[0, 3], ph * 16 / 1000, rtol=1e-5)  # wh_ratio_clip
    p8 = core.thetaobb2pointobb(gt)
    d8 = core.pointobb2delta(prop, p8)
    assert torch.allclose(core.delta2pointobb(prop, d8), p8, atol=1e-3)
    r = gt.clone()
    out = core.thetaobb_rescale(r, 2.0, reverse_flag=True)
    assert out is r and torch.allclose(r[:, :4], gt[:, :4] / 2) and torch.equal(r[:, 4], gt[:, 4])     # in place, theta kept
    res = core.rbbox2result(torch.zeros(0, 6), torch.zeros(0), 16)
    assert len(res) == 15 and res[0].shape == (0, 6)


def test_hobb_codecs():
    # transforms.py:137-163 written out for one box: first edge (10,20)->(50,20), height 30 -> a = pi/2, corners below it
    # (x4 = 10 - 30 cos(pi/2) = 9.999999999999998 in double, which the reference's int() truncates to 9, :161)
    assert core.hobb2pointobb([10, 20, 50, 20, 30]) == [10, 20, 50, 20, 50, 50, 9, 50]
    t = torch.tensor([[10., 20., 50., 20., 30.], [0., 0., 30., 40., 10.]])
    p = core.hobb2pointobb(t)
    assert torch.allclose(p[0], torch.tensor([10., 20., 50., 20., 50., 50., 10., 50.]), atol=1e-4)
    e = torch.tensor([30., 40.]) / 50.0                       # second box: unit edge (0.6, 0.8); p3 = p2 + h * (-0.8, 0.6)
    assert torch.allclose(p[1, 4:6], torch.tensor([30. - 8., 40. + 6.]), atol=1e-4)
    assert torch.allclose(p[1, 6:8], torch.tensor([-8., 6.]), atol=1e-4)
    side = (p[1, 4:6] - p[1, 2:4])
    assert abs(float(side @ e)) < 1e-4 and abs(float(side.norm()) - 10) < 1e-4     # perpendicular, length h
    prop = torch.tensor([[5., 10., 60., 70.]]).repeat(2, 1)
    d = core.hobb2delta(prop, t, stds=(0.1, 0.1, 0.1, 0.1, 0.2))
    back = core.delta2hobb(prop, d, stds=(0.1, 0.1, 0.1, 0.1, 0.2))
    assert torch.allclose(back[:, :4], t[:, :4], atol=1e-3) and torch.allclose(back[:, 4], t[:, 4] + 1, atol=1e-3)   # gh = h + 1 (:539)
    r = t.clone()
    assert core.hobb_rescale(r, 2.0, reverse_flag=True) is r and torch.allclose(r, t / 2)


@pytest.mark.gpu
def test_get_det_rbboxes_matches_class_loop(cuda):
    """decode + rescale + one batched launch == the reference-shaped per-class loop with the single-class op."""
    from aidet_b200 import synth
    from aidet_b200.ops import thetaobb_nms
    g = torch.Generator().manual_seed(3)
    n, C = 600, 15
    boxes, _ = synth.dota_boxes(n, side=1024, seed=5)
    hb = core.pointobb2bbox(core.thetaobb2pointobb(boxes))
    rois = torch.cat([torch.zeros(n, 1), hb], 1)
    gt = boxes[:, None, :].repeat(1, C + 1, 1) + torch.randn(n, C + 1, 5, generator=g) * torch.tensor([2., 2., 2., 2., 0.02])
    gt[..., 2:4] = gt[..., 2:4].clamp(min=4)
    stds = (0.1, 0.1, 0.2, 0.2, 0.1)
    pred = torch.stack([core.thetaobb2delta(hb, gt[:, c], stds=stds) for c in range(C + 1)], 1).reshape(n, -1)
    logits = torch.randn(n, C + 1, generator=g) * 3
    cfg = dict(score_thr=0.05, polygon_nms_iou_thr=0.5, max_per_img=200)
    dets, labels = core.get_det_rbboxes(rois.to(cuda), logits.to(cuda), pred.to(cuda), (1024, 1024, 3), 2.0, rescale=True,
                                        cfg=cfg, target_stds=stds)
    assert dets.shape[1] == 6 and dets.shape[0] == labels.shape[0] <= 200
    # reference-shaped loop (rbbox_nms.py:29-49 with thetaobb_nms in the slot of rbbox_nms.py:97)
    rb, sc = core.get_det_rbboxes(rois.to(cuda), logits.to(cuda), pred.to(cuda), (1024, 1024, 3), 2.0, rescale=True, cfg=None,
                                  target_stds=stds)
    rb = rb.view(n, C + 1, 5)
    ref_d, ref_l = [], []
    for c in range(1, C + 1):
        m = sc[:, c] > 0.05
        if not m.any():
            continue
        cls = torch.cat([rb[m, c], sc[m, c, None]], 1)
        kept, _ = thetaobb_nms(cls, 0.5)
        ref_d.append(kept); ref_l.append(torch.full((kept.size(0),), c - 1, dtype=torch.long, device=cuda))
    ref_d, ref_l = torch.cat(ref_d), torch.cat(ref_l)
    if ref_d.size(0) > 200:
        _, inds = ref_d[:, -1].sort(descending=True)
        ref_d, ref_l = ref_d[inds[:200]], ref_l[inds[:200]]
    assert torch.equal(dets, ref_d) and torch.equal(labels, ref_l)
