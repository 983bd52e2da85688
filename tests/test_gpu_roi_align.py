"""GPU parity: (rotated) RoIAlign forward / backward vs the float64 oracle and torchvision CPU.

Tolerance (BASELINE.json north_star): 1e-4 relative.  Written here as
|got - ref| <= 1e-4 * |ref| + 1e-4 * rms(ref) element-wise.
"""
import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.ops import RoIAlign, RoIAlignRotated, roi_align, roi_align_rotated
from aidet_b200.ops import functional as F
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _close(got, ref, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    rms = float(np.sqrt(np.mean(ref ** 2))) if ref.size else 0.0
    err = np.abs(got - ref) - (RTOL * np.abs(ref) + RTOL * rms)
    assert not np.isnan(got).any(), what + ": NaN in output"
    assert (err <= 0).all(), "%s: max violation %.3g (rms %.3g)" % (what, err.max(), rms)


def _hbb_rois(k, n, side, seed):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(k, generator=g) * side * 0.8
    y1 = torch.rand(k, generator=g) * side * 0.8
    w = torch.rand(k, generator=g) * side * 0.5 + 1
    h = torch.rand(k, generator=g) * side * 0.5 + 1
    b = torch.randint(0, n, (k,), generator=g).float()
    return torch.stack([b, x1, y1, x1 + w, y1 + h], 1)


@pytest.mark.parametrize("aligned", [False, True])
@pytest.mark.parametrize("sample_num", [0, 2])
def test_axis_aligned_vs_oracle_and_torchvision(cuda, aligned, sample_num):
    """gradcheck.py:11-30 shapes: feat 2x16x15x15, 20 rois, out 3, scale 1/8 -- plus a larger case."""
    from torchvision.ops import roi_align as tv_roi_align
    for (n, c, hw, k, out, scale, side) in [(2, 16, 15, 20, 3, 1 / 8, 120), (2, 32, 50, 64, 7, 1 / 4, 200)]:
        g = torch.Generator().manual_seed(hw)
        feat = torch.randn(n, c, hw, hw, generator=g)
        rois = _hbb_rois(k, n, side, seed=hw + 1)
        mod = RoIAlign(out, scale, sample_num, aligned=aligned)
        got = mod(feat.to(cuda), rois.to(cuda))
        assert got.shape == (k, c, out, out)
        variant = O.ROI_V2_ALIGNED if aligned else O.ROI_V1
        ref = O.roi_align_fwd(feat.permute(0, 2, 3, 1).numpy(), rois.numpy(), scale, (out, out), sample_num, variant)
        _close(got.permute(0, 2, 3, 1).cpu().numpy(), ref, "fwd vs oracle")
        if not aligned:      # v1 == torchvision with x2+1, y2+1
            r2 = rois.clone()
            r2[:, 3:] += 1
            tv = tv_roi_align(feat.double(), r2.double(), (out, out), scale, sample_num, aligned=False)
        else:
            tv = tv_roi_align(feat.double(), rois.double(), (out, out), scale, sample_num, aligned=True)
        # torchvision (aligned=False) clamps the RoI to >= 1 feature px, the reference's v1 kernel does not
        # (roi_align_kernel.cu:85-86 clamps at 0): the two agree exactly only for RoIs >= 1 px
        big = (((rois[:, 3] + 1) * scale - rois[:, 1] * scale >= 1) & ((rois[:, 4] + 1) * scale - rois[:, 2] * scale >= 1))
        assert aligned or big.sum() >= k // 2
        sel = torch.ones(k, dtype=torch.bool) if aligned else big
        _close(got.cpu().numpy()[sel.numpy()], tv.numpy()[sel.numpy()], "fwd vs torchvision")


@pytest.mark.parametrize("aligned", [False, True])
def test_rotated_theta0_equals_axis_aligned(cuda, aligned):
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 32, 40, 40, generator=g).to(cuda)
    rois5 = _hbb_rois(50, 2, 150, seed=4)
    x1, y1, x2, y2 = rois5[:, 1], rois5[:, 2], rois5[:, 3], rois5[:, 4]
    rois6 = torch.stack([rois5[:, 0], (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1, torch.zeros(50)], 1)
    a = RoIAlign(7, 0.25, 2, aligned=aligned)(feat, rois5.to(cuda))
    b = RoIAlignRotated(7, 0.25, 2, aligned=aligned)(feat, rois6.to(cuda))
    _close(b.cpu().numpy(), a.cpu().numpy(), "theta=0")


@pytest.mark.parametrize("variant_aligned", [False, True])
@pytest.mark.parametrize("sample_num", [0, 2])
def test_rotated_forward_backward_vs_oracle(cuda, variant_aligned, sample_num):
    g = torch.Generator().manual_seed(7)
    n, c, hw, k, out, scale = 2, 64, 48, 96, 7, 0.125
    feat = torch.randn(n, c, hw, hw, generator=g)
    rois, _ = synth.rotated_rois(k // n, n, tile=int(hw / scale), seed=5)
    rois[:, 3:5] = rois[:, 3:5].clamp(max=200)
    f = feat.to(cuda).requires_grad_(True)
    y = roi_align_rotated(f, rois.to(cuda), (out, out), scale, sample_num, variant_aligned)
    variant = O.ROI_V2_ALIGNED if variant_aligned else O.ROI_V1
    ref = O.roi_align_fwd(feat.permute(0, 2, 3, 1).numpy(), rois.numpy(), scale, (out, out), sample_num, variant)
    _close(y.detach().permute(0, 2, 3, 1).cpu().numpy(), ref, "rotated fwd")
    go = torch.randn(y.shape, generator=g)
    y.backward(go.to(cuda))
    gref = O.roi_align_bwd(go.permute(0, 2, 3, 1).contiguous().numpy(), (n, hw, hw, c), rois.numpy(), scale,
                           sample_num, variant)
    _close(f.grad.permute(0, 2, 3, 1).cpu().numpy(), gref, "rotated bwd")


def test_axis_aligned_backward_vs_oracle(cuda):
    g = torch.Generator().manual_seed(9)
    n, c, hw, k, out, scale = 2, 16, 15, 20, 3, 1 / 8
    feat = torch.randn(n, c, hw, hw, generator=g)
    rois = _hbb_rois(k, n, 120, seed=10)
    for aligned in (False, True):
        f = feat.to(cuda).requires_grad_(True)
        y = roi_align(f, rois.to(cuda), out, scale, 2, aligned)
        go = torch.randn(y.shape, generator=g)
        res = torch.autograd.grad(y, f, go.to(cuda))[0]
        variant = O.ROI_V2_ALIGNED if aligned else O.ROI_V1
        gref = O.roi_align_bwd(go.permute(0, 2, 3, 1).contiguous().numpy(), (n, hw, hw, c), rois.numpy(), scale, 2,
                               variant)
        _close(res.permute(0, 2, 3, 1).cpu().numpy(), gref, "bwd aligned=%s" % aligned)


def test_autograd_contract(cuda):
    """grad to features only; rois and scalars get None (roi_align.py:60-73); CPU raises (roi_align.py:41-42)."""
    feat = torch.randn(1, 8, 10, 10, device=cuda, requires_grad=True)
    rois = torch.tensor([[0, 20.0, 20.0, 30.0, 16.0, 0.4]], device=cuda, requires_grad=True)
    y = roi_align_rotated(feat, rois, 3, 0.25, 2, True)
    y.sum().backward()
    assert feat.grad is not None and rois.grad is None
    with pytest.raises(NotImplementedError):
        roi_align_rotated(torch.randn(1, 8, 10, 10), torch.zeros(1, 6), 3, 0.25, 2, True)
    with pytest.raises(AssertionError):
        RoIAlignRotated(3, 0.25)(feat, torch.zeros(1, 5, device=cuda))
    m = RoIAlign(7, 1 / 4, 2)
    assert m.out_size == (7, 7) and "RoIAlign(out_size=(7, 7)" in repr(m)


def test_odd_channels_and_channels_last(cuda):
    g = torch.Generator().manual_seed(12)
    feat = torch.randn(2, 13, 20, 20, generator=g)          # C % 4 != 0 -> scalar lanes
    rois, _ = synth.rotated_rois(16, 2, tile=80, seed=6)
    y = roi_align_rotated(feat.to(cuda), rois.to(cuda), 5, 0.25, 2, True)
    ref = O.roi_align_fwd(feat.permute(0, 2, 3, 1).numpy(), rois.numpy(), 0.25, (5, 5), 2, O.ROI_V2_ALIGNED)
    _close(y.permute(0, 2, 3, 1).cpu().numpy(), ref, "C=13")
    feat2 = torch.randn(2, 64, 20, 20, generator=g)
    y_a = roi_align_rotated(feat2.to(cuda), rois.to(cuda), 5, 0.25, 2, True)
    y_b = roi_align_rotated(feat2.to(cuda).contiguous(memory_format=torch.channels_last), rois.to(cuda), 5, 0.25, 2,
                            True)
    assert torch.equal(y_a, y_b)
    empty = roi_align_rotated(feat2.to(cuda), torch.zeros((0, 6), device=cuda), 5, 0.25, 2, True)
    assert empty.shape == (0, 64, 5, 5)


def test_border_rules(cuda):
    """RoIs hanging over every image edge (rejection at [-1,H], clamps at 0 and H-1, roi_align_kernel.cu:22-46)."""
    g = torch.Generator().manual_seed(13)
    feat = torch.randn(1, 8, 12, 12, generator=g)
    rois = torch.tensor([[0, -10.0, -10.0, 40.0, 30.0, 0.3], [0, 50.0, 50.0, 60.0, 20.0, -1.0],
                         [0, 24.0, -30.0, 10.0, 80.0, 0.0], [0, 24.0, 24.0, 200.0, 200.0, 0.7],
                         [0, 47.9, 47.9, 3.0, 3.0, 0.1]])
    for aligned in (False, True):
        for sn in (0, 2):
            y = roi_align_rotated(feat.to(cuda), rois.to(cuda), 4, 0.25, sn, aligned)
            variant = O.ROI_V2_ALIGNED if aligned else O.ROI_V1
            ref = O.roi_align_fwd(feat.permute(0, 2, 3, 1).numpy(), rois.numpy(), 0.25, (4, 4), sn, variant)
            _close(y.permute(0, 2, 3, 1).cpu().numpy(), ref, "border aligned=%s sn=%d" % (aligned, sn))


def test_c3_multilevel_small(cuda):
    """Config C3 shape in miniature: 4 FPN levels in ONE launch through the functional API."""
    feats = synth.fpn_features(batch=2, channels=64, tile=256, seed=3)
    rois, lvl = synth.rotated_rois(64, 2, tile=256, seed=3)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    out = F.rroi_align_forward([f.to(cuda) for f in feats], rois.to(cuda), scales, (7, 7), 2, 2, lvl.to(cuda))
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    grads = [torch.zeros_like(f, device=cuda) for f in feats]
    F.rroi_align_backward(go.to(cuda), grads, rois.to(cuda), scales, 2, 2, lvl.to(cuda))
    for l in range(4):
        sel = (lvl == l).nonzero().flatten()
        if sel.numel() == 0:
            continue
        ref = O.roi_align_fwd(feats[l].numpy(), rois[sel].numpy(), scales[l], (7, 7), 2, O.ROI_V2_ALIGNED)
        _close(out[sel.to(cuda)].cpu().numpy(), ref, "level %d fwd" % l)
        gref = O.roi_align_bwd(go[sel].contiguous().numpy(), tuple(feats[l].shape), rois[sel].numpy(), scales[l], 2,
                               O.ROI_V2_ALIGNED)
        _close(grads[l].cpu().numpy(), gref, "level %d bwd" % l)
    # deterministic gather backward: overwrites (buffers start as garbage), equals the oracle, bit-reproducible
    g1 = [torch.full_like(f, float("nan"), device=cuda) for f in feats]
    g2 = [torch.full_like(f, 7.0, device=cuda) for f in feats]
    g3 = [torch.full_like(f, -3.0, device=cuda) for f in feats]
    F.rroi_align_backward_gather(go.to(cuda), g1, rois.to(cuda), scales, 2, 2, lvl.to(cuda), deterministic=True)
    F.rroi_align_backward_gather(go.to(cuda), g2, rois.to(cuda), scales, 2, 2, lvl.to(cuda), deterministic=True)
    F.rroi_align_backward_gather(go.to(cuda), g3, rois.to(cuda), scales, 2, 2, lvl.to(cuda), deterministic=False)
    for l in range(4):
        assert torch.equal(g1[l], g2[l])
        assert torch.allclose(g1[l], g3[l], rtol=1e-5, atol=1e-5)
        sel = (lvl == l).nonzero().flatten()
        gref = O.roi_align_bwd(go[sel].contiguous().numpy(), tuple(feats[l].shape), rois[sel].numpy(), scales[l], 2,
                               O.ROI_V2_ALIGNED) if sel.numel() else np.zeros(tuple(feats[l].shape))
        _close(g1[l].cpu().numpy(), gref, "level %d gather bwd" % l)


def test_gather_backward_edge_cases(cuda):
    """no RoIs -> zeros; RoIs outside the image / bad batch index contribute nothing; C = 8 and C = 40 lanes."""
    for c in (8, 40):
        g = torch.Generator().manual_seed(c)
        gf = [torch.full((2, 12, 12, c), float("nan"), device=cuda)]
        F.rroi_align_backward_gather(torch.zeros((0, 3, 3, c), device=cuda), gf, torch.zeros((0, 6), device=cuda), [0.25], 2, 2)
        assert (gf[0] == 0).all()
        rois = torch.tensor([[0, 20.0, 20.0, 30.0, 16.0, 0.4], [1, 500.0, 500.0, 20.0, 20.0, 0.1],
                             [5, 20.0, 20.0, 10.0, 10.0, 0.0], [1, -4.0, 10.0, 30.0, 30.0, 1.2]])
        go = torch.randn(4, 3, 3, c, generator=g)
        for det in (False, True):
            gf = [torch.full((2, 12, 12, c), float("nan"), device=cuda)]
            F.rroi_align_backward_gather(go.to(cuda), gf, rois.to(cuda), [0.25], 2, 2, deterministic=det)
        ok = rois[:, 0] < 2
        gref = O.roi_align_bwd(go[ok].contiguous().numpy(), (2, 12, 12, c), rois[ok].numpy(), 0.25, 2, O.ROI_V2_ALIGNED)
        _close(gf[0].cpu().numpy(), gref, "gather edge C=%d" % c)


@pytest.mark.parametrize("c,out,sample_num", [(256, 7, 2), (192, 14, 2), (512, 7, 1), (40, 14, 1), (1024, 3, 2), (1280, 3, 2),
                                              (64, 9, 2)])
def test_taplist_forward_shapes(cuda, c, out, sample_num):
    """The tap-list forward over its whole shape range: 1-8 channel chunks per bin (C = 40 ... 1024; 1280 falls back to
    the per-sample kernel), one and several CTAs per RoI (7x7 = 49 bins, 14x14 = 196 bins), sample_num 1 and 2, RoIs
    hanging over the border, fully outside, and with a batch index out of range (-> zeros)."""
    g = torch.Generator().manual_seed(c + out)
    n, hw, scale = 2, 40, 0.25
    feat = torch.randn(n, hw, hw, c, generator=g)                       # NHWC
    rois, _ = synth.rotated_rois(24, n, tile=int(hw / scale), seed=c)
    rois[:, 3:5] = rois[:, 3:5].clamp(max=120)
    extra = torch.tensor([[0, -6.0, 80.0, 40.0, 30.0, 0.5], [1, 400.0, 400.0, 20.0, 20.0, 0.2], [7, 50.0, 50.0, 30.0, 30.0, 0.1],
                          [1, 159.5, 159.5, 6.0, 2.0, -0.9], [0, 80.0, 80.0, 1.0, 1.0, 0.0]])
    rois = torch.cat([rois, extra])
    for variant in (O.ROI_V1, O.ROI_V2_ALIGNED):
        y = F.rroi_align_forward([feat.to(cuda)], rois.to(cuda), [scale], (out, out), sample_num, variant)
        ok = rois[:, 0] < n
        ref = O.roi_align_fwd(feat.numpy(), rois[ok].numpy(), scale, (out, out), sample_num, variant)
        _close(y[ok.to(cuda)].cpu().numpy(), ref, "taplist C=%d out=%d sn=%d v=%d" % (c, out, sample_num, variant))
        assert float(y[~ok.to(cuda)].abs().max()) == 0.0
