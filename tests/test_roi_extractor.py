"""SingleRoIExtractor mirror (mmdet/models/roi_extractors/single_level.py:11-107): level mapping, rescale and the
reference's per-level loop on CPU (through torchvision, which the reference itself offers: roi_align.py:138-141);
on the GPU the fused one-launch path against that loop, forward and backward."""
import math

import pytest
import torch

from aidet_b200.models import SingleRoIExtractor


def _rois(k, seed, side=1024):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(k, 2, generator=g) * side
    wh = torch.exp(torch.rand(k, 2, generator=g) * (math.log(600) - math.log(8)) + math.log(8))
    b = torch.randint(0, 2, (k, 1), generator=g).float()
    return torch.cat([b, c - wh / 2, c + wh / 2], 1)


def test_map_roi_levels_and_rescale():
    ext = SingleRoIExtractor(dict(type='RoIAlign', out_size=7, sample_num=2), 16, [4, 8, 16, 32])
    assert ext.num_inputs == 4 and ext.roi_layers[2].spatial_scale == 1 / 16 and ext.roi_layers[0].out_size == (7, 7)
    rois = _rois(500, 0)
    lv = ext.map_roi_levels(rois, 4)
    for r, l in zip(rois.tolist(), lv.tolist()):              # single_level.py:57-60
        scale = math.sqrt((r[3] - r[1] + 1) * (r[4] - r[2] + 1))
        want = 0 if scale < 112 else 1 if scale < 224 else 2 if scale < 448 else 3
        if min(abs(scale - t) for t in (112, 224, 448)) > 1e-3:
            assert l == want
    assert lv.dtype == torch.long and lv.min() >= 0 and lv.max() <= 3
    r2 = ext.roi_rescale(rois, 1.5)                           # same centre, sides (+1 model) x 1.5
    assert torch.allclose((r2[:, 1] + r2[:, 3]) / 2, (rois[:, 1] + rois[:, 3]) / 2, atol=1e-3)
    assert torch.allclose(r2[:, 3] - r2[:, 1] + 1, (rois[:, 3] - rois[:, 1] + 1) * 1.5, rtol=1e-5)
    # rotated rois: level by sqrt((w+1)(h+1))
    rext = SingleRoIExtractor(dict(type='RoIAlignRotated', out_size=7, sample_num=2), 16, [4, 8, 16, 32])
    rr = torch.tensor([[0, 50., 50., 55., 55., 0.3], [0, 50., 50., 223., 223., 1.0]])
    assert rext.map_roi_levels(rr, 4).tolist() == [0, 2]
    with pytest.raises(AssertionError):
        SingleRoIExtractor(dict(type='NoSuchLayer', out_size=7), 16, [4])


def test_reference_loop_on_cpu_torchvision():
    tv = pytest.importorskip("torchvision")
    ext = SingleRoIExtractor(dict(type='RoIAlign', out_size=5, sample_num=2, use_torchvision=True), 8, [4, 8, 16, 32])
    assert not ext._fusable()
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(2, 8, 256 // s, 256 // s, generator=g) for s in (1, 2, 4, 8)]
    rois = _rois(40, 2)
    out = ext(feats, rois)
    lv = ext.map_roi_levels(rois, 4)
    for i in range(4):
        m = lv == i
        if m.any():
            ref = tv.ops.roi_align(feats[i], rois[m], (5, 5), 1 / (4 * 2 ** i), 2)
            assert torch.equal(out[m], ref)
    one = ext(feats[:1], rois)                                # single level: straight through (single_level.py:90-91)
    assert torch.equal(one, tv.ops.roi_align(feats[0], rois, (5, 5), 0.25, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("layer,aligned", [("RoIAlign", False), ("RoIAlign", True), ("RoIAlignRotated", True)])
def test_fused_extractor_matches_level_loop(cuda, layer, aligned):
    ext = SingleRoIExtractor(dict(type=layer, out_size=7, sample_num=2, aligned=aligned), 32, [4, 8, 16, 32]).to(cuda)
    assert ext._fusable()
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(2, 32, 512 // s, 512 // s, generator=g).to(cuda).requires_grad_(True) for s in (4, 8, 16, 32)]
    rois = _rois(300, 6, side=512)
    if layer == "RoIAlignRotated":
        c = (rois[:, 1:3] + rois[:, 3:5]) / 2
        wh = rois[:, 3:5] - rois[:, 1:3]
        th = (torch.rand(rois.size(0), 1, generator=g) - 0.5) * math.pi
        rois = torch.cat([rois[:, :1], c, wh, th], 1)
    rois = rois.to(cuda)
    out = ext(feats, rois)
    go = torch.randn(out.shape, generator=g).to(cuda)
    out.backward(go)
    fused_grads = [f.grad.clone() for f in feats]
    for f in feats:
        f.grad = None
    # the reference's loop over levels with the single-level op (single_level.py:96-107)
    lv = ext.map_roi_levels(rois, 4)
    ref = out.new_zeros(out.shape)
    for i in range(4):
        m = lv == i
        if m.any():
            ref[m] = ext.roi_layers[i](feats[i], rois[m])
    assert torch.equal(out, ref)
    ref.backward(go)
    for fg, f in zip(fused_grads, feats):
        rg = f.grad if f.grad is not None else torch.zeros_like(fg)      # a level no RoI maps to gets no gradient
        tol = 1e-4 * (rg.abs().max().item() + 1e-6)
        assert (fg - rg).abs().max().item() <= tol            # summation order differs (taps bucketed per launch)
    assert (lv.bincount(minlength=4) > 0).sum() >= 3          # several levels were exercised
