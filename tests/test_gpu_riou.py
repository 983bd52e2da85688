"""GPU parity: rotated IoU (C ABI -> aidet_b200.core.rbbox_overlaps) vs the float64 oracle.

Tolerance (BASELINE.json north_star): IoU within 1e-5 absolute.
"""
import math

import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.core import rbbox_overlaps
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _check(a, b, cuda, mode="iou", tol=TOL):
    got = rbbox_overlaps(a.to(cuda), b.to(cuda), mode=mode).cpu().numpy().astype(np.float64)
    ref = O.riou_matrix(a.numpy(), b.numpy(), mode=mode)
    err = np.abs(got - ref)
    assert err.max() <= tol, "max err %.3g at %s" % (err.max(), np.unravel_index(err.argmax(), err.shape))
    return got, ref


@pytest.mark.parametrize("dense", [False, True])
def test_c1_matrix_2000(cuda, dense):
    """Config C1: 2000 x 2000 theta-OBB matrix (seeds 0 / 1), DOTA-shaped and dense variants."""
    a, _ = synth.dota_boxes(2000, seed=0, dense=dense)
    b, _ = synth.dota_boxes(2000, seed=1, dense=dense)
    got, ref = _check(a, b, cuda)
    if dense:
        assert (ref > 0).all()


def test_iof_mode(cuda):
    a, _ = synth.dota_boxes(700, seed=5)
    b, _ = synth.dota_boxes(900, seed=6)
    _check(a, b, cuda, mode="iof")


def test_pointobb_matches_thetaobb(cuda):
    a, _ = synth.dota_boxes(600, seed=7)
    b, _ = synth.dota_boxes(500, seed=8)
    a8, b8 = synth.thetaobb2pointobb(a), synth.thetaobb2pointobb(b)
    got8, _ = _check(a8, b8, cuda)
    got5, _ = _check(a, b, cuda)
    assert np.abs(got8 - got5).max() <= 2e-5          # 8-point inputs are rounded to f32


def test_general_quads(cuda):
    """point-OBB heads regress free quadrilaterals: convex but not rectangles."""
    a, _ = synth.dota_boxes(500, side=300, seed=9)
    b, _ = synth.dota_boxes(500, side=300, seed=10)
    a8, ca = synth.free_quads(a, 0.1, seed=1)
    b8, cb = synth.free_quads(b, 0.1, seed=2)
    a8, b8 = a8[ca].contiguous(), b8[cb].contiguous()
    assert len(a8) > 400 and len(b8) > 400
    # clockwise input order must give the same answer
    b8_cw = b8.view(-1, 4, 2).flip(1).reshape(-1, 8).contiguous()
    got, ref = _check(a8, b8, cuda)
    got_cw, _ = _check(a8, b8_cw, cuda)
    assert np.abs(got - got_cw).max() <= 5e-6
    _check(a8, b8, cuda, mode="iof", tol=3e-5)       # IoF of point-OBBs: 3e-5 (accuracy notes)
    # simple non-convex quads follow the signed-triangle (polyiou lineage) definition
    dart = torch.tensor([[0, 0, 10, 0, 3, 3, 0, 10.0]])
    sq = torch.tensor([[1, 1, 9, 1, 9, 9, 1, 9.0], [0, 0, 4, 0, 4, 4, 0, 4], [5, 5, 12, 5, 12, 12, 5, 12]])
    for x, y in ((dart, sq), (sq, dart)):
        g = rbbox_overlaps(x.to(cuda), y.to(cuda)).cpu().numpy()
        assert np.abs(g - O.riou_matrix(x.numpy(), y.numpy(), algo=O.ALGO_FAN)).max() < 2e-6


def test_known_answers(cuda):
    sq = torch.tensor([[0, 0, 2, 2, 0.0]])
    rot = torch.tensor([[0, 0, 2, 2, math.pi / 4]])
    v = rbbox_overlaps(sq.to(cuda), rot.to(cuda)).item()
    assert abs(v - 0.70710678) < 1e-6                  # inter = 8(sqrt2 - 1)
    far = torch.tensor([[100, 100, 2, 2, 0.3]])
    assert rbbox_overlaps(sq.to(cuda), far.to(cuda)).item() == 0.0
    box = torch.tensor([[10, 10, 4, 2, 0.3]])
    for other in ([10, 10, 4, 2, 0.3], [10, 10, 4, 2, 0.3 + math.pi], [10, 10, 2, 4, 0.3 + math.pi / 2]):
        v = rbbox_overlaps(box.to(cuda), torch.tensor([other], dtype=torch.float32).to(cuda)).item()
        assert abs(v - 1.0) < 2e-6
    # theta = 0 pair: closed-form axis-aligned IoU without +1
    a = torch.tensor([[5, 5, 10, 10, 0.0]])
    b = torch.tensor([[10, 5, 10, 10, 0.0]])
    assert abs(rbbox_overlaps(a.to(cuda), b.to(cuda)).item() - 50.0 / 150.0) < 1e-6


def test_symmetry_and_invariance(cuda):
    a, _ = synth.dota_boxes(300, side=200, seed=12)
    b, _ = synth.dota_boxes(300, side=200, seed=13)
    m_ab = rbbox_overlaps(a.to(cuda), b.to(cuda))
    m_ba = rbbox_overlaps(b.to(cuda), a.to(cuda))
    assert (m_ab - m_ba.t()).abs().max().item() <= 2e-6
    shift = torch.tensor([500.0, -300.0, 0, 0, 0])
    m_sh = rbbox_overlaps((a + shift).to(cuda), (b + shift).to(cuda))
    assert (m_ab - m_sh).abs().max().item() <= 5e-6


def test_aligned_and_empty(cuda):
    a, _ = synth.dota_boxes(1000, seed=14)
    b = a.clone()
    b[:, :2] += 3.0
    b[:, 4] += 0.1
    got = rbbox_overlaps(a.to(cuda), b.to(cuda), is_aligned=True).cpu().numpy()
    ref = O.riou_aligned(a.numpy(), b.numpy())
    assert got.shape == (1000,) and np.abs(got - ref).max() <= TOL
    got_f = rbbox_overlaps(a.to(cuda), b.to(cuda), mode="iof", is_aligned=True).cpu().numpy()
    assert np.abs(got_f - O.riou_aligned(a.numpy(), b.numpy(), mode="iof")).max() <= TOL
    e = torch.zeros((0, 5), device=cuda)
    assert tuple(rbbox_overlaps(e, a[:1].to(cuda)).shape) == (0, 1)
    assert tuple(rbbox_overlaps(a[:1].to(cuda), e).shape) == (1, 0)
    assert tuple(rbbox_overlaps(e, e).shape) == (0, 0)
    assert tuple(rbbox_overlaps(e, e, is_aligned=True).shape) == (0, 1)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("fmt", [5, 8])
def test_ragged_sizes(cuda, dense, fmt):
    """tile tails: sizes around the 256-column / 64-row tile edges, odd row counts for the two-row steps; the dense sets
    take the solid-tile loop (no per-pair circle test), the others the dual-row / early-out loops; both modes."""
    for m, n in [(1, 1), (2, 3), (3, 256), (63, 257), (65, 255), (130, 513), (1, 1000), (1000, 1), (129, 1)]:
        a, _ = synth.dota_boxes(m, side=200, seed=20 + m, dense=dense)
        b, _ = synth.dota_boxes(n, side=200, seed=40 + n, dense=dense)
        if fmt == 8:
            a, b = synth.thetaobb2pointobb(a).float(), synth.thetaobb2pointobb(b).float()
        _check(a, b, cuda)
        _check(a, b, cuda, mode="iof", tol=3e-5 if fmt == 8 else TOL)


def test_solid_and_sparse_tiles_in_one_matrix(cuda):
    """A column set that mixes a dense cluster with far-away boxes: warps of one tile take different loops (solid /
    dual-row / early-out); far pairs must come out exactly 0 where the bounding circles are disjoint on the non-solid paths
    and within tolerance everywhere."""
    a, _ = synth.dota_boxes(700, side=16384, seed=71, dense=True)
    b1, _ = synth.dota_boxes(512, side=16384, seed=72, dense=True)
    b2, _ = synth.dota_boxes(700, side=16384, seed=73, dense=False)
    b = torch.cat([b1[:256], b2[:300], b1[256:], b2[300:]])
    got, ref = _check(a, b, cuda)
    assert (got[:, 256:556][ref[:, 256:556] == 0] < 1e-6).all()


def test_cpu_tensor_raises():
    a, _ = synth.dota_boxes(4, seed=1)
    with pytest.raises(NotImplementedError):
        rbbox_overlaps(a, a)


@pytest.mark.gpu
def test_hbb_bbox_overlaps_reference_doctest_and_oracle(cuda):
    """bbox_overlaps (mmdet/core/bbox/geometry.py:4-88) served by the tiled kernel, fmt 4: the doctest matrix of
    geometry.py:22-44, the empty-shape contracts of :54-55, random boxes against the oracle restatement, iof, aligned."""
    from aidet_b200.core import bbox_overlaps
    b1 = torch.tensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]], dtype=torch.float32, device=cuda)
    b2 = torch.tensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]], dtype=torch.float32, device=cuda)
    got = bbox_overlaps(b1, b2)
    want = torch.tensor([[0.5238, 0.0500, 0.0041], [0.0323, 0.0452, 1.0000], [0.0000, 0.0000, 0.0000]])
    assert torch.allclose(got.cpu(), want, atol=1e-4)
    empty = torch.empty(0, 4, device=cuda)
    nonempty = torch.tensor([[0., 0., 10., 9.]], device=cuda)
    assert tuple(bbox_overlaps(empty, nonempty).shape) == (0, 1)
    assert tuple(bbox_overlaps(nonempty, empty).shape) == (1, 0)
    assert tuple(bbox_overlaps(empty, empty).shape) == (0, 0)
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(700, 2, generator=g) * 300
    a = torch.cat([xy, xy + torch.rand(700, 2, generator=g) * 80], 1)
    xy = torch.rand(533, 2, generator=g) * 300
    b = torch.cat([xy, xy + torch.rand(533, 2, generator=g) * 80], 1)
    for mode in ("iou", "iof"):
        ref = O.hbb_overlaps(a.numpy(), b.numpy(), mode=mode, plus_one=True)
        got = bbox_overlaps(a.to(cuda), b.to(cuda), mode=mode).cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-5
        al = bbox_overlaps(a[:533].to(cuda), b.to(cuda), mode=mode, is_aligned=True).cpu().numpy()
        assert np.abs(al - np.diag(O.hbb_overlaps(a[:533].numpy(), b.numpy(), mode=mode, plus_one=True))).max() <= 1e-5


@pytest.mark.parametrize("m,n,fmt", [(300, 1028, 5), (33, 256, 5), (1, 4, 5), (700, 2000, 8), (257, 1030, 5), (64, 1023, 5)])
def test_multi_destination_stores(cuda, m, n, fmt):
    """aidet_riou_matrix_multi_f32 (the fused compute + all-gather kernel of the row-sharded form) with several LOCAL
    destinations: every destination must hold exactly what the single-destination kernel writes.  n % 4 == 0 takes the
    TMA tensor-store kernel (tiles clipped at the matrix edge by the tensor map -- in 16-byte units, hence the
    condition), the other shapes the per-value peer-store kernel; rows past m and the padding columns of a wider
    buffer must stay untouched either way."""
    from aidet_b200.ops import functional as F
    a, _ = synth.dota_boxes(m, side=400, seed=m)
    b, _ = synth.dota_boxes(n, side=400, seed=n + 1)
    if fmt == 8:
        a, b = synth.thetaobb2pointobb(a).float(), synth.thetaobb2pointobb(b).float()
    a, b = a.to(cuda), b.to(cuda)
    want = F.riou_matrix(a, b)
    ld = (n + 7) // 4 * 4 + 4                                        # a wider buffer: row stride > n, still 16-byte rows
    bufs = [torch.full((m + 3, ld), -7.0, device=cuda) for _ in range(3)]
    F.riou_matrix_multi(a, b, [t.data_ptr() for t in bufs], ld)
    torch.cuda.synchronize()
    for t in bufs:
        assert torch.equal(t[:m, :n], want)
        assert bool((t[m:] == -7.0).all()) and bool((t[:, n:] == -7.0).all())
    if n % 4:                                                        # odd row stride: falls back to the per-value stores
        bufs = [torch.full((m, n), -7.0, device=cuda) for _ in range(2)]
        F.riou_matrix_multi(a, b, [t.data_ptr() for t in bufs], n)
        for t in bufs:
            assert torch.equal(t, want)
