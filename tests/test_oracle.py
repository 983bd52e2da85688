"""CPU: pin the oracle (test infrastructure) against every golden vector the reference holds for
this path, analytic known answers, cv2, torchvision and the reference's own
nms_cpu.cpp compiled unmodified (oracle/_ref)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- box conventions
def test_thetaobb2pointobb_fixtures():
    """theta-OBB fixtures of tests/test_randomflip.py:6-7 -> cv2.boxPoints (transforms.py:45-55)."""
    boxes = [[200, 200, 300, 150, 45 * math.pi / 180.0], [700, 800, 300, 200, 135 * math.pi / 180.0]]
    want = np.array([[40.901, 146.967, 146.967, 40.901, 359.099, 253.033, 253.033, 359.099],
                     [735.355, 623.223, 876.777, 764.645, 664.645, 976.777, 523.223, 835.355]])
    got = O.thetaobb2pointobb(boxes)
    assert np.abs(got - want).max() < 1e-3


def test_thetaobb2pointobb_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    b = np.stack([rng.uniform(0, 1000, 50), rng.uniform(0, 1000, 50), rng.uniform(2, 300, 50),
                  rng.uniform(2, 300, 50), rng.uniform(-math.pi, math.pi, 50)], 1).astype(np.float32)
    got = O.thetaobb2pointobb(b)
    for i in range(50):
        ref = cv2.boxPoints(((float(b[i, 0]), float(b[i, 1])), (float(b[i, 2]), float(b[i, 3])),
                             float(b[i, 4]) * 180.0 / np.pi)).reshape(-1)
        assert np.abs(got[i] - ref).max() < 2e-3       # cv2 works in float32


# ---------------------------------------------------------------- rotated IoU
def test_riou_known_answers():
    sq = [[0, 0, 2, 2, 0.0]]
    for algo in (O.ALGO_SH, O.ALGO_FAN):
        assert abs(O.riou_matrix(sq, [[0, 0, 2, 2, math.pi / 4]], algo=algo)[0, 0] - 0.70710678118) < 1e-9
        assert O.riou_matrix(sq, [[10, 10, 2, 2, 0.3]], algo=algo)[0, 0] == 0.0
        assert abs(O.riou_matrix(sq, sq, algo=algo)[0, 0] - 1.0) < 1e-12
        assert abs(O.riou_matrix([[5, 5, 10, 10, 0]], [[10, 5, 10, 10, 0]], algo=algo)[0, 0] - 1 / 3) < 1e-12
        # w=h box vs itself rotated 90 degrees; theta == theta + pi; (w,h,theta) == (h,w,theta+pi/2)
        assert abs(O.riou_matrix([[3, 4, 6, 6, 0.2]], [[3, 4, 6, 6, 0.2 + math.pi / 2]], algo=algo)[0, 0] - 1) < 1e-6
        assert abs(O.riou_matrix([[3, 4, 8, 2, 0.2]], [[3, 4, 8, 2, 0.2 + math.pi]], algo=algo)[0, 0] - 1) < 1e-6
        assert abs(O.riou_matrix([[3, 4, 8, 2, 0.2]], [[3, 4, 2, 8, 0.2 + math.pi / 2]], algo=algo)[0, 0] - 1) < 1e-6
    # iof = inter / area of the first box
    assert abs(O.riou_matrix([[0, 0, 2, 2, 0]], [[0, 0, 10, 10, 0]], mode="iof")[0, 0] - 1.0) < 1e-12
    assert abs(O.riou_matrix([[0, 0, 10, 10, 0]], [[0, 0, 2, 2, 0]], mode="iof")[0, 0] - 0.04) < 1e-12


def test_riou_two_algorithms_agree():
    """Sutherland-Hodgman vs the DOTA_devkit-lineage triangle fan, random DOTA-shaped boxes."""
    from aidet_b200 import synth
    a, _ = synth.dota_boxes(400, side=400, seed=3)
    b, _ = synth.dota_boxes(400, side=400, seed=4)
    m0 = O.riou_matrix(a.numpy(), b.numpy(), algo=O.ALGO_SH)
    m1 = O.riou_matrix(a.numpy(), b.numpy(), algo=O.ALGO_FAN)
    assert (m0 > 0).mean() > 0.02 and np.abs(m0 - m1).max() < 1e-7
    assert np.abs(m0 - O.riou_matrix(b.numpy(), a.numpy()).T).max() < 1e-12
    a8, b8 = synth.thetaobb2pointobb(a), synth.thetaobb2pointobb(b)
    assert np.abs(O.riou_matrix(a8.numpy(), b8.numpy()) - m0).max() < 2e-5     # 8-point inputs rounded to f32


def test_riou_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    from aidet_b200 import synth
    a, _ = synth.dota_boxes(120, side=300, seed=5)
    m = O.riou_matrix(a.numpy(), a.numpy())
    a = a.numpy()
    worst = 0.0
    for i in range(120):
        for j in range(0, 120, 7):
            r1 = ((float(a[i, 0]), float(a[i, 1])), (float(a[i, 2]), float(a[i, 3])), float(np.degrees(a[i, 4])))
            r2 = ((float(a[j, 0]), float(a[j, 1])), (float(a[j, 2]), float(a[j, 3])), float(np.degrees(a[j, 4])))
            ret, pts = cv2.rotatedRectangleIntersection(r1, r2)
            ar = cv2.contourArea(cv2.convexHull(pts)) if ret > 0 and pts is not None and len(pts) > 2 else 0.0
            iou = ar / (a[i, 2] * a[i, 3] + a[j, 2] * a[j, 3] - ar)
            worst = max(worst, abs(iou - m[i, j]))
    assert worst < 1e-3          # cv2 is float32 and loose on near-degenerate cases


# ---------------------------------------------------------------- HBB overlaps / NMS
def test_bbox_overlaps_doctest_vector():
    """mmdet/core/bbox/geometry.py:22-37 doctest."""
    b1 = [[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]]
    b2 = [[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]]
    want = np.array([[0.5238, 0.0500, 0.0041], [0.0323, 0.0452, 1.0000], [0.0, 0.0, 0.0]])
    assert np.abs(O.hbb_overlaps(b1, b2) - want).max() < 5e-5


NMS4 = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9],
                 [35.3, 11.5, 39.9, 14.5, 0.4], [35.2, 11.7, 39.7, 15.7, 0.3]], dtype=np.float32)
NMS7 = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9], [49.2, 31.8, 51.0, 35.4, 0.5],
                 [35.1, 11.5, 39.1, 15.7, 0.5], [35.6, 11.8, 39.3, 14.2, 0.5], [35.3, 11.5, 39.9, 14.5, 0.4],
                 [35.2, 11.7, 39.7, 15.7, 0.3]], dtype=np.float32)


def test_hbb_nms_reference_vectors():
    """tests/test_nms.py:16-41 (4 boxes @0.7 -> 3 kept) and nms_wrapper.py:25-34 (7 boxes -> 3)."""
    for cmp_ge in (True, False):
        keep, _ = O.nms(NMS4[:, :4], NMS4[:, 4], 0.7, cmp_ge=cmp_ge, plus_one=True)
        assert list(keep) == [0, 2, 3]
        keep, _ = O.nms(NMS7[:, :4], NMS7[:, 4], 0.7, cmp_ge=cmp_ge, plus_one=True)
        assert len(keep) == 3


def test_hbb_nms_matches_compiled_reference():
    """oracle_nms (fmt 4, +1, >=) == the reference's nms_cpu.cpp (oracle/_ref), on random boxes."""
    build_ref.build()
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    assert ref.nms(torch.from_numpy(NMS4), 0.7).tolist() == [0, 2, 3]
    rng = np.random.default_rng(1)
    for n, thr in [(300, 0.3), (800, 0.5), (500, 0.7)]:
        xy = rng.uniform(0, 200, (n, 2))
        wh = rng.uniform(5, 80, (n, 2))
        dets = np.concatenate([xy, xy + wh, rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
        want = ref.nms(torch.from_numpy(dets), thr).numpy()
        got, near = O.nms(dets[:, :4], dets[:, 4], thr, cmp_ge=True, plus_one=True, margin=1e-6)
        if near == 0:        # the reference computes in float32, the oracle in float64
            assert np.array_equal(got, want)
        else:
            bad, _ = O.nms_verify(dets[:, :4], dets[:, 4], thr, want, cmp_ge=True, plus_one=True, margin=1e-5)
            assert bad == 0


def test_rotated_nms_semantics():
    boxes = np.array([[10, 10, 8, 4, 0.1], [10.5, 10, 8, 4, 0.1], [30, 30, 8, 4, 1.0], [10, 10, 8, 4, 0.1 + math.pi]],
                     dtype=np.float32)
    scores = np.array([0.5, 0.9, 0.8, 0.9], dtype=np.float32)
    keep, _ = O.nms(boxes, scores, 0.5)
    assert list(keep) == [1, 2]              # tie 1/3 broken by index; ascending original index out
    keep, _ = O.nms(boxes, scores, 0.5, groups=[0, 1, 0, 2])
    assert list(keep) == [0, 1, 2, 3]
    # > vs >= at an exactly representable IoU (two unit-offset squares: 1/3 is not, use 0.5: 2x2 vs shift 2/3)
    b2 = np.array([[0, 0, 4, 4, 0], [0, 2, 4, 4, 0]], dtype=np.float32)       # inter 8, union 24 -> 1/3
    iou = O.riou_matrix(b2[:1], b2[1:])[0, 0]
    k_gt, _ = O.nms(b2, [0.9, 0.8], iou, cmp_ge=False)
    k_ge, _ = O.nms(b2, [0.9, 0.8], iou, cmp_ge=True)
    assert list(k_gt) == [0, 1] and list(k_ge) == [0]
    bad, near = O.nms_verify(b2, [0.9, 0.8], iou, [0, 1])
    assert bad == 0 and near == 1
    bad, _ = O.nms_verify(boxes, scores, 0.5, [0, 1, 2])
    assert bad > 0
    assert len(O.nms(np.zeros((0, 5), np.float32), np.zeros((0,), np.float32), 0.5)[0]) == 0


# ---------------------------------------------------------------- RoIAlign
def _feat_rois(seed, n=2, c=8, hw=15, k=20, side=120):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(n, c, hw, hw, generator=g, dtype=torch.float64)
    x1 = torch.rand(k, generator=g, dtype=torch.float64) * side * 0.8
    y1 = torch.rand(k, generator=g, dtype=torch.float64) * side * 0.8
    w = torch.rand(k, generator=g, dtype=torch.float64) * side * 0.5 + 1
    h = torch.rand(k, generator=g, dtype=torch.float64) * side * 0.5 + 1
    b = torch.randint(0, n, (k,), generator=g).double()
    rois = torch.stack([b, x1, y1, x1 + w, y1 + h], 1)
    return feat.float().double(), rois.float().double()


@pytest.mark.parametrize("sample_num", [0, 2])
def test_roi_align_v1_v2_vs_torchvision(sample_num):
    """v1 == torchvision(rois with x2+1,y2+1, aligned=False); v2 == torchvision(aligned=...)."""
    tv = pytest.importorskip("torchvision.ops")
    feat, rois = _feat_rois(0)
    nhwc = feat.permute(0, 2, 3, 1).numpy()
    # v1 clamps the RoI size at 0 (roi_align_kernel.cu:85-86), torchvision at 1: equal for RoIs >= 1 feature px
    big = ((rois[:, 3] + 1 - rois[:, 1]) >= 8) & ((rois[:, 4] + 1 - rois[:, 2]) >= 8)
    assert big.sum() >= 10
    r1 = rois[big].clone()
    r1[:, 3:] += 1
    want = tv.roi_align(feat, r1, (3, 3), 1 / 8, sample_num, aligned=False).permute(0, 2, 3, 1).numpy()
    got = O.roi_align_fwd(nhwc, rois[big].numpy(), 1 / 8, (3, 3), sample_num, O.ROI_V1)
    assert np.abs(got - want).max() < 1e-6            # rois pass through float32 in the oracle's C ABI
    want = tv.roi_align(feat, rois, (3, 3), 1 / 8, sample_num, aligned=True).permute(0, 2, 3, 1).numpy()
    got = O.roi_align_fwd(nhwc, rois.numpy(), 1 / 8, (3, 3), sample_num, O.ROI_V2_ALIGNED)
    assert np.abs(got - want).max() < 1e-6
    if sample_num > 0:
        want = tv.roi_align(feat, rois, (3, 3), 1 / 8, sample_num, aligned=False).permute(0, 2, 3, 1).numpy()
        got = O.roi_align_fwd(nhwc, rois.numpy(), 1 / 8, (3, 3), sample_num, O.ROI_V2)
        assert np.abs(got - want).max() < 1e-6


def test_roi_align_backward_vs_torchvision_autograd():
    tv = pytest.importorskip("torchvision.ops")
    feat, rois = _feat_rois(1)
    feat.requires_grad_(True)
    y = tv.roi_align(feat, rois, (3, 3), 1 / 8, 2, aligned=True)
    go = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).double()
    y.backward(go)
    got = O.roi_align_bwd(go.permute(0, 2, 3, 1).contiguous().numpy(), (2, 15, 15, 8), rois.numpy(), 1 / 8, 2,
                          O.ROI_V2_ALIGNED)
    assert np.abs(got - feat.grad.permute(0, 2, 3, 1).numpy()).max() < 1e-5


@pytest.mark.parametrize("variant", [O.ROI_V1, O.ROI_V2, O.ROI_V2_ALIGNED])
def test_rotated_roi_align_reduces_to_axis_aligned(variant):
    feat, rois5 = _feat_rois(3)
    nhwc = feat.permute(0, 2, 3, 1).numpy()
    if variant == O.ROI_V2:
        # aligned=False clamps the RoI to >= 1 px: anchored at (x1,y1) for HBB RoIs, centred for rotated
        # ones, so the two only coincide for RoIs of at least one feature pixel
        rois5 = rois5[((rois5[:, 3] - rois5[:, 1]) >= 8) & ((rois5[:, 4] - rois5[:, 2]) >= 8)]
    x1, y1, x2, y2 = rois5[:, 1], rois5[:, 2], rois5[:, 3], rois5[:, 4]
    rois6 = torch.stack([rois5[:, 0], (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1, torch.zeros_like(x1)], 1)
    a = O.roi_align_fwd(nhwc, rois5.numpy(), 1 / 8, (3, 3), 2, variant)
    b = O.roi_align_fwd(nhwc, rois6.numpy(), 1 / 8, (3, 3), 2, variant)
    assert np.abs(a - b).max() < 1e-5
    # a 90-degree rotation of a square RoI samples the transposed grid
    sq = np.array([[0, 60.0, 60.0, 40.0, 40.0, 0.0]], np.float32)
    sq90 = sq.copy()
    sq90[0, 5] = math.pi / 2
    o0 = O.roi_align_fwd(nhwc, sq, 1 / 8, (3, 3), 2, variant)[0]
    o90 = O.roi_align_fwd(nhwc, sq90, 1 / 8, (3, 3), 2, variant)[0]
    assert np.abs(o90 - np.flip(o0.transpose(1, 0, 2), axis=0)).max() < 1e-5      # o90[p,q] = o0[q, n-1-p]


def test_map_roi_levels():
    """single_level.py:54-73."""
    rois = np.array([[0, 0, 0, 10, 10], [0, 0, 0, 110, 110], [0, 0, 0, 111, 111], [0, 0, 0, 300, 300],
                     [0, 0, 0, 2000, 2000]], np.float32)
    assert list(O.map_roi_levels(rois)) == [0, 0, 1, 2, 3]


def test_golden_files_match_oracle():
    """tests/golden/*.npz were produced by tests/golden/make_golden.py (cv2 + torchvision + the compiled
    reference nms_cpu) -- the oracle must reproduce them."""
    g = np.load(os.path.join(GOLD, "golden_v1.npz"))
    assert np.abs(O.thetaobb2pointobb(g["theta_boxes"]) - g["theta_points_cv2"]).max() < 2e-3
    assert np.abs(O.riou_matrix(g["theta_boxes"], g["theta_boxes"]) - g["riou_cv2"]).max() < 1e-3
    keep, near = O.nms(g["hbb_dets"][:, :4], g["hbb_dets"][:, 4], 0.5, cmp_ge=True, plus_one=True)
    assert near == 0 and np.array_equal(keep, g["hbb_keep_ref_0p5"])
    nhwc = np.ascontiguousarray(g["feat_nchw"].transpose(0, 2, 3, 1))
    for name, variant in (("v1", O.ROI_V1), ("v2a", O.ROI_V2_ALIGNED)):
        got = O.roi_align_fwd(nhwc, g["rois5"], 0.125, (3, 3), 2, variant)
        assert np.abs(got - g["roi_fwd_" + name].transpose(0, 2, 3, 1)).max() < 1e-6
    got = O.roi_align_bwd(np.ascontiguousarray(g["roi_grad_out"].transpose(0, 2, 3, 1)), nhwc.shape, g["rois5"],
                          0.125, 2, O.ROI_V2_ALIGNED)
    assert np.abs(got - g["roi_bwd_v2a"].transpose(0, 2, 3, 1)).max() < 1e-5


def test_float32_roialign_noise_floor():
    """Why the RoIAlign parity bound carries a coordinate-resolution term (tests/test_gpu_full_configs.py): torchvision's
    OWN float32 CPU kernel -- the implementation the reference offers beside its CUDA op, roi_align.py:138-141 -- deviates
    from its float64 self by a few ulp32(W) * rms(features) on a 256-px map, far outside 1e-4 relative element-wise,
    while it is well inside `1e-4 |ref| + 4 ulp32(W) rms`.  No GPU, no aidet_b200 code involved."""
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(1, 16, 256, 256, generator=g)
    c = torch.rand(400, 2, generator=g) * 900 + 60
    wh = torch.rand(400, 2, generator=g) * 100 + 20
    rois = torch.cat([torch.zeros(400, 1), c - wh / 2, c + wh / 2], 1)
    lo = tv.ops.roi_align(feat, rois, (7, 7), 0.25, 2, aligned=True).double().numpy()
    hi = tv.ops.roi_align(feat.double(), rois.double(), (7, 7), 0.25, 2, aligned=True).numpy()
    err = np.abs(lo - hi)
    floor = 4.0 * float(np.spacing(np.float32(256.0))) * float(feat.std())
    assert (err > 1e-4 * np.abs(hi)).mean() > 1e-3            # pure element-wise 1e-4 relative is not attainable in float32
    assert (err <= 1e-4 * np.abs(hi) + floor).all()           # the bound the GPU tests use holds for it
    assert err.max() > 0.02 * floor                           # and is not slack by orders of magnitude


def test_greedy_block_fixed_point_equals_sequential_greedy():
    """csrc/rnms.cu: greedy_block solves the 32 greedy decisions of a diagonal block by fixed-point iteration
    K <- ~(init | OR_{j in K} D_j) over an upper-triangular bit matrix (D_j only holds bits > j).  Restated here in numpy
    and compared with the sequential greedy loop on random, dense, empty and chain-shaped (worst case: 32 rounds) blocks."""
    rng = np.random.default_rng(12)

    def sequential(init, D):
        cur, keep = int(init), 0
        for k in range(32):
            if not (cur >> k) & 1:
                keep |= 1 << k
                cur |= int(D[k])
        return keep

    def fixed_point(init, D):
        K = ~int(init) & 0xffffffff
        for rounds in range(1, 40):
            R = int(init)
            for j in range(32):
                if (K >> j) & 1:
                    R |= int(D[j])
            Kn = ~R & 0xffffffff
            if Kn == K:
                return K, rounds
            K = Kn
        raise AssertionError("no fixed point within 39 rounds")

    upper = np.array([(0xffffffff << (j + 1)) & 0xffffffff for j in range(32)], dtype=np.uint64)
    cases = []
    for density in (0.0, 0.02, 0.1, 0.5, 1.0):
        for _ in range(400):
            bits = rng.random((32, 32)) < density
            D = np.array([sum(1 << c for c in range(32) if bits[j, c]) for j in range(32)], dtype=np.uint64) & upper
            cases.append((int(rng.integers(0, 1 << 32)) if rng.random() < 0.5 else 0, D))
    chain = np.array([(1 << (j + 1)) & 0xffffffff for j in range(32)], dtype=np.uint64)      # box j suppresses only box j + 1
    cases.append((0, chain))
    worst = 0
    for init, D in cases:
        keep, rounds = fixed_point(init, D)
        assert keep == sequential(init, D)
        worst = max(worst, rounds)
    assert worst <= 33
