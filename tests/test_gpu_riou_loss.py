"""GPU: rotated IoU loss (aidet_riou_aligned_grad_f32 through RotatedIoULoss / rotated_iou) against central
differences of the float64 oracle overlap, plus the reduction / weight contract of iou_loss.py:129-165."""
import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.models import RotatedIoULoss, riou_loss, rotated_iou
from aidet_b200.ops import functional as F
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["iou", "iof"])
def test_gradient_vs_finite_differences(cuda, mode):
    pred, target = synth.regression_pairs(20000, seed=11)
    ov, ga, gb = F.riou_aligned_grad(pred.to(cuda), target.to(cuda), None, mode)
    ref, fd = O.riou_aligned_grad_fd(pred.numpy(), target.numpy(), mode, 1e-5)
    _, fd2 = O.riou_aligned_grad_fd(pred.numpy(), target.numpy(), mode, 2e-5)
    smooth = np.abs(fd - fd2).max(1) < 1e-6                  # pairs on a kink of the overlap are excluded
    assert smooth.mean() > 0.99
    assert np.abs(ov.cpu().numpy() - ref).max() <= 1e-5      # BASELINE.json: IoU within 1e-5 absolute
    g = torch.cat([ga, gb], 1).cpu().numpy().astype(np.float64)
    err = np.abs(g - fd)[smooth]
    assert err.max() < 2e-5, err.max()
    # forward of the gradient kernel == the aligned kernel
    assert float((ov - F.riou_aligned(pred.to(cuda), target.to(cuda), mode)).abs().max()) <= 1e-6


def test_loss_backward_and_reductions(cuda):
    pred, target = synth.regression_pairs(3000, seed=12)
    p = pred.to(cuda).requires_grad_(True)
    t = target.to(cuda)
    loss = riou_loss(p, t, reduction='none')
    ref, fd = O.riou_aligned_grad_fd(pred.numpy(), target.numpy(), "iou", 1e-5)
    _, fd2 = O.riou_aligned_grad_fd(pred.numpy(), target.numpy(), "iou", 2e-5)
    smooth = (np.abs(fd - fd2).max(1) < 1e-6) & (ref > 1e-3)
    want = -np.log(np.maximum(ref, 1e-6))
    assert np.abs(loss.detach().cpu().numpy() - want)[ref > 1e-3].max() < 1e-3
    w = torch.rand(3000, device=cuda)
    (loss * w).sum().backward()
    want_g = -(w.cpu().numpy() / np.maximum(ref, 1e-6))[:, None] * fd[:, :5]
    got = p.grad.cpu().numpy()
    rel = np.abs(got - want_g)[smooth] / (np.abs(want_g)[smooth].max() + 1e-12)
    assert np.abs(got - want_g)[smooth].max() < 1e-3 * max(1.0, np.abs(want_g)[smooth].max()), rel.max()
    # pairs below eps get no gradient (clamp), disjoint pairs stay finite
    far = t.clone()
    far[:, 0] += 5000
    p2 = pred.to(cuda).requires_grad_(True)
    l2 = riou_loss(p2, far, reduction='sum')
    l2.backward()
    assert torch.isfinite(l2) and float(p2.grad.abs().max()) == 0.0
    assert abs(float(l2.detach()) / 3000 + np.log(1e-6)) < 1e-4

    m = RotatedIoULoss(loss_weight=2.0)
    base = riou_loss(p.detach(), t, reduction='none')
    assert torch.allclose(m(p.detach(), t), 2.0 * base.mean())
    assert torch.allclose(m(p.detach(), t, weight=w, avg_factor=100.0), 2.0 * (base * w).sum() / 100.0)
    assert torch.allclose(m(p.detach(), t, reduction_override='sum'), 2.0 * base.sum())
    assert m(p.detach(), t, reduction_override='none').shape == (3000,)
    assert float(m(p, t, weight=torch.zeros(3000, 1, device=cuda))) == 0.0


def test_target_gradient_and_descent(cuda):
    """Both inputs get gradients; gradient descent on the loss increases the IoU (end-to-end sanity of the sign)."""
    pred, target = synth.regression_pairs(512, seed=13)
    p = pred.to(cuda).requires_grad_(True)
    t = target.to(cuda).requires_grad_(True)
    iou0 = rotated_iou(p, t)
    iou0.sum().backward()
    assert p.grad is not None and t.grad is not None
    assert float((p.grad[:, :2] + t.grad[:, :2]).abs().max()) < 1e-4          # translation invariance
    # gradient ASCENT on the IoU with steps scaled to the box (d iou / d px ~ 1 / size): the mean IoU must rise
    x = pred.to(cuda).clone().requires_grad_(True)
    tt = target.to(cuda)
    first = float(rotated_iou(x, tt).mean())
    for _ in range(40):
        x.grad = None
        rotated_iou(x, tt).sum().backward()
        with torch.no_grad():
            s2 = (x[:, 2] * x[:, 3]).unsqueeze(1)
            x[:, :4] += 0.05 * s2 * x.grad[:, :4]
            x[:, 4] += 0.05 * x.grad[:, 4]
    assert float(rotated_iou(x, tt).mean()) > first + 0.25


@pytest.mark.parametrize("mode", ["iou", "iof"])
def test_point_obb_gradient_and_loss(cuda, mode):
    """fmt 8: gradient w.r.t. the corner coordinates of convex quads vs central differences of the float64 oracle,
    and the loss on (n, 8) boxes through autograd."""
    pred, target = synth.regression_pairs(8000, seed=15)
    a8, b8 = synth.thetaobb2pointobb(pred), synth.thetaobb2pointobb(target)
    g = torch.Generator().manual_seed(3)
    size = torch.sqrt(pred[:, 2] * pred[:, 3])[:, None]
    a8 = (a8 + torch.randn(a8.shape, generator=g) * 0.03 * size).contiguous()
    b8 = (b8 + torch.randn(b8.shape, generator=g) * 0.03 * size).contiguous()
    a8[::2] = a8[::2].reshape(-1, 4, 2).flip(1).reshape(-1, 8)
    ov, ga, gb = F.riou_aligned_grad(a8.to(cuda), b8.to(cuda), None, mode)
    ref, fd = O.riou_aligned_grad_fd(a8.numpy(), b8.numpy(), mode, 1e-5)
    _, fd2 = O.riou_aligned_grad_fd(a8.numpy(), b8.numpy(), mode, 2e-5)
    smooth = np.abs(fd - fd2).max(1) < 1e-6
    assert smooth.mean() > 0.99 and np.abs(ov.cpu().numpy() - ref).max() <= 1e-5
    got = torch.cat([ga, gb], 1).cpu().numpy().astype(np.float64)
    assert np.abs(got - fd)[smooth].max() < 2e-5
    if mode == "iou":
        p = a8.to(cuda).requires_grad_(True)
        loss = riou_loss(p, b8.to(cuda), reduction='sum')
        loss.backward()
        want = -fd[:, :8] / np.maximum(ref, 1e-6)[:, None]
        ok = smooth & (ref > 1e-3)
        assert np.abs(p.grad.cpu().numpy() - want)[ok].max() < 1e-3 * max(1.0, np.abs(want[ok]).max())
