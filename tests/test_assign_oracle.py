"""CPU: the numpy restatement of MaxIoUAssigner (oracle/oracle.py) against outputs of the REFERENCE's own class
(tests/golden/golden_assign_v1.npz, made by tests/golden/make_golden_assign.py from
/root/reference/mmdet/core/bbox/assigners/max_iou_assigner.py), the analytic rotated-IoU gradient of csrc/geom.cuh
(compiled for the host) against central differences of the float64 oracle, and the host-side classes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from aidet_b200 import synth
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_assign_v1.npz"))
N_CFG = GOLD["configs"].shape[0]


def golden_cfg(ci):
    pos, nlo, nhi, mp, allg, ign, wrt, is_pair = GOLD["configs"][ci]
    neg = (float(nlo), float(nhi)) if is_pair else float(nhi)
    return float(pos), neg, float(mp), bool(allg), float(ign), bool(wrt)


@pytest.mark.parametrize("ci", range(N_CFG))
def test_oracle_assign_wrt_overlaps_matches_reference(ci):
    pos, neg, mp, allg, _, _ = golden_cfg(ci)
    gi, mo, lb = O.max_iou_assign_wrt_overlaps(GOLD["c%d_mat" % ci], pos, neg, mp, allg, GOLD["c%d_gt_labels" % ci][:17])
    assert np.array_equal(gi, GOLD["c%d_mat_gt_inds" % ci])
    assert np.array_equal(mo, GOLD["c%d_mat_max_overlaps" % ci])
    assert np.array_equal(lb, GOLD["c%d_mat_labels" % ci])
    # and on the reference's own float32 overlap matrix of the box case (no ignore boxes in this form)
    if golden_cfg(ci)[4] <= 0:
        gi, mo, lb = O.max_iou_assign_wrt_overlaps(GOLD["c%d_overlaps" % ci], pos, neg, mp, allg, GOLD["c%d_gt_labels" % ci])
        assert np.array_equal(gi, GOLD["c%d_gt_inds" % ci]) and np.array_equal(lb, GOLD["c%d_labels" % ci])


@pytest.mark.parametrize("ci", range(N_CFG))
def test_oracle_assign_on_boxes_matches_reference(ci):
    pos, neg, mp, allg, ign, wrt = golden_cfg(ci)
    gi, mo, lb = O.max_iou_assign(GOLD["c%d_boxes" % ci], GOLD["c%d_gts" % ci], pos, neg, mp, allg, ign, wrt,
                                  GOLD["c%d_ign" % ci], GOLD["c%d_gt_labels" % ci])
    assert np.array_equal(gi, GOLD["c%d_gt_inds" % ci])
    assert np.abs(mo - GOLD["c%d_max_overlaps" % ci]).max() < 1e-6
    assert np.array_equal(lb, GOLD["c%d_labels" % ci])


def test_oracle_assign_doctest_and_empty():
    # max_iou_assigner.py:78-84
    gi, _, _ = O.max_iou_assign(np.array([[0, 0, 10, 10], [10, 10, 20, 20]], np.float32),
                                np.array([[0, 0, 10, 9]], np.float32), 0.5, 0.5)
    assert gi.tolist() == [1, 0]
    gi, mo, lb = O.max_iou_assign_wrt_overlaps(np.zeros((0, 4), np.float32), 0.5, 0.5, gt_labels=np.zeros((0,), np.int64))
    assert gi.tolist() == [0, 0, 0, 0] and mo.tolist() == [0, 0, 0, 0] and lb.tolist() == [0, 0, 0, 0]


def test_host_classes_empty_and_cpu_contract():
    from aidet_b200.core import AssignResult, MaxIoUAssigner
    a = MaxIoUAssigner(0.5, 0.5)
    r = a.assign(torch.zeros(4, 5), torch.zeros(0, 5), gt_labels=torch.zeros(0, dtype=torch.long))
    assert isinstance(r, AssignResult) and r.num_gts == 0 and r.gt_inds.tolist() == [0] * 4 and r.labels.tolist() == [0] * 4
    r = a.assign(torch.zeros(0, 5), torch.zeros(3, 5))
    assert r.num_preds == 0 and r.labels is None
    with pytest.raises(NotImplementedError):
        a.assign(torch.rand(4, 5), torch.rand(2, 5))
    r = AssignResult(2, torch.tensor([0, 2, -1]), torch.tensor([0.1, 0.8, 0.3]), torch.tensor([0, 7, 0]))
    r.add_gt_(torch.tensor([5, 7]))
    assert r.gt_inds.tolist() == [1, 2, 0, 2, -1] and r.labels.tolist() == [5, 7, 0, 7, 0]
    assert r.max_overlaps.tolist()[:2] == [1.0, 1.0] and "num_gts=2" in repr(r)


def test_loss_module_contract_on_cpu():
    from aidet_b200.models import RotatedIoULoss, riou_loss, rotated_iou
    m = RotatedIoULoss(loss_weight=2.0)
    pred = torch.rand(3, 5, requires_grad=True)
    z = m(pred, torch.rand(3, 5), weight=torch.zeros(3, 1))        # all weights 0 -> (pred * weight).sum(), iou_loss.py:145-146
    assert float(z.detach()) == 0.0 and z.requires_grad
    with pytest.raises(NotImplementedError):
        riou_loss(pred, torch.rand(3, 5))
    assert rotated_iou(torch.zeros(0, 5), torch.zeros(0, 5)).shape == (0,)


@pytest.fixture(scope="module")
def sim_grad():
    src = os.path.join(HERE, "hostsim", "geom_sim.cpp")
    hdr = os.path.join(HERE, "..", "aidet_b200", "csrc", "geom.cuh")
    so = os.path.join(HERE, "hostsim", "libgeom_sim.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=fast", "-mfma", "-o", so, src])
    lib = C.CDLL(so)

    def run(a, b, mode):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.empty((len(a),), np.float32)
        g = np.empty((len(a), 10), np.float32)
        fn = lib.sim_riou_aligned_grad if a.shape[1] == 5 else lib.sim_riou_aligned_grad8
        g = np.empty((len(a), 2 * a.shape[1]), np.float32)
        fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(a), mode, out.ctypes.data_as(C.c_void_p),
           g.ctypes.data_as(C.c_void_p))
        return out.astype(np.float64), g.astype(np.float64)
    return run


@pytest.mark.parametrize("mode,mi", [("iou", 0), ("iof", 1)])
def test_analytic_riou_gradient_matches_finite_differences(sim_grad, mode, mi):
    """csrc/geom.cuh: rect_overlap_grad (boundary-integral gradient, FP32) vs central differences of the float64
    oracle overlap.  Pairs where two step sizes disagree sit on a kink of the overlap and are excluded."""
    a, b = synth.regression_pairs(4000, seed=3)
    a, b = a.numpy(), b.numpy()
    ov, g = sim_grad(a, b, mi)
    ref, fd = O.riou_aligned_grad_fd(a, b, mode, 1e-5)
    _, fd2 = O.riou_aligned_grad_fd(a, b, mode, 2e-5)
    smooth = np.abs(fd - fd2).max(1) < 1e-6
    assert smooth.mean() > 0.99
    assert np.abs(ov - ref).max() < 1e-5
    assert (ref > 0.05).mean() > 0.8                     # the set really overlaps
    err = np.abs(g - fd)[smooth]
    assert err.max() < 2e-5, err.max()
    # translation invariance: d/d centre of a == -d/d centre of b
    assert np.abs(g[:, 0] + g[:, 5]).max() < 1e-5 and np.abs(g[:, 1] + g[:, 6]).max() < 1e-5


def test_analytic_riou_gradient_special_cases(sim_grad):
    # disjoint -> 0; b inside a: IoU = area_b / area_a, d/dw_a = -area_b/(w_a^2 h_a), no dependence on centres / angles
    a = np.array([[0, 0, 10, 10, 0.3], [0, 0, 20, 10, 0.2], [0, 0, -20, 10, 0.2]], np.float32)
    b = np.array([[100, 0, 10, 10, 0.1], [1, 0.5, 4, 2, 0.9], [1, 0.5, 4, 2, 0.9]], np.float32)
    ov, g = sim_grad(a, b, 0)
    assert ov[0] == 0 and np.abs(g[0]).max() == 0
    assert abs(ov[1] - 8 / 200) < 1e-6
    want = np.array([0, 0, -8 / (20 * 20 * 10), -8 / (20 * 10 * 10), 0, 0, 0, 2 / 200, 4 / 200, 0])
    assert np.abs(g[1] - want).max() < 1e-6
    want[2] = -want[2]                                   # negative w: |w| is used, the gradient flips sign
    assert np.abs(g[2] - want).max() < 1e-6


@pytest.mark.parametrize("mode,mi", [("iou", 0), ("iof", 1)])
@pytest.mark.parametrize("noise", [0.0, 0.04])
def test_point_obb_gradient_matches_finite_differences(sim_grad, mode, mi, noise):
    """csrc/geom.cuh: quad_overlap_grad (gradient w.r.t. the 2 x 8 corner coordinates of convex quads, FP32) vs central
    differences of the float64 oracle; rectangles and perturbed (general convex) quads, both orientations."""
    pred, target = synth.regression_pairs(3000, seed=5)
    a8, b8 = synth.thetaobb2pointobb(pred).numpy(), synth.thetaobb2pointobb(target).numpy()
    rng = np.random.default_rng(2)
    size = torch.sqrt(pred[:, 2] * pred[:, 3]).numpy()[:, None]
    a8 = (a8 + rng.normal(0, noise, a8.shape) * size).astype(np.float32)
    b8 = (b8 + rng.normal(0, noise, b8.shape) * size).astype(np.float32)
    a8[::2] = a8[::2].reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8)          # clockwise corner order for half of them
    b8[::3] = b8[::3].reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8)
    ov, g = sim_grad(a8, b8, mi)
    ref, fd = O.riou_aligned_grad_fd(a8, b8, mode, 1e-5)
    _, fd2 = O.riou_aligned_grad_fd(a8, b8, mode, 2e-5)
    smooth = np.abs(fd - fd2).max(1) < 1e-6
    assert smooth.mean() > 0.99 and np.abs(ov - ref).max() < 1e-5 and (ref > 0.05).mean() > 0.9
    assert np.abs(g - fd)[smooth].max() < 2e-5
    # translating both quads together changes nothing: the 16 x- (y-) derivatives sum to 0
    assert np.abs(g[:, 0::2].sum(1)).max() < 1e-5 and np.abs(g[:, 1::2].sum(1)).max() < 1e-5


# ---- the reference's own assigner tests (tests/test_assigner.py:17-162), golden vectors included
REF_BBOXES = [[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]]
REF_GTS = [[0, 0, 10, 9], [0, 10, 10, 19]]


def test_reference_assigner_goldens_on_the_oracle():
    b, g = np.array(REF_BBOXES, np.float32), np.array(REF_GTS, np.float32)
    gi, _, lb = O.max_iou_assign(b, g, 0.5, 0.5, gt_labels=np.array([2, 3]))
    assert gi.tolist() == [1, 0, 2, 0] and len(lb) == 4                              # test_assigner.py:17-38
    gi, _, _ = O.max_iou_assign(b, g, 0.5, 0.5, ignore_iof_thr=0.5, ignore_wrt_candidates=False,
                                gt_bboxes_ignore=np.array([[30, 30, 40, 40]], np.float32))
    assert gi.tolist() == [1, 0, 2, -1]                                              # test_assigner.py:41-65


def test_reference_assigner_empty_cases_on_the_mirror():
    """tests/test_assigner.py:68-162: no truths, no boxes, neither -- decided on the host, no device needed."""
    from aidet_b200.core import MaxIoUAssigner
    a = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5)
    bboxes = torch.FloatTensor(REF_BBOXES)
    r = a.assign(bboxes, torch.FloatTensor([]))                                      # :68-86
    assert torch.all(r.gt_inds == torch.LongTensor([0, 0, 0, 0]))
    gts, labels = torch.FloatTensor(REF_GTS), torch.LongTensor([2, 3])
    r = a.assign(torch.empty((0, 4)), gts, gt_labels=labels)                         # :89-112
    assert len(r.gt_inds) == 0 and tuple(r.labels.shape) == (0,)
    r = a.assign(torch.empty((0, 4)), gts, gt_labels=None)
    assert len(r.gt_inds) == 0 and r.labels is None
    ai = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5, ignore_iof_thr=0.5)        # :115-148
    ign = torch.Tensor([[30, 30, 40, 40]])
    r = ai.assign(torch.empty((0, 4)), gts, gt_labels=labels, gt_bboxes_ignore=ign)
    assert len(r.gt_inds) == 0 and tuple(r.labels.shape) == (0,)
    r = ai.assign(torch.empty((0, 4)), gts, gt_labels=None, gt_bboxes_ignore=ign)
    assert len(r.gt_inds) == 0 and r.labels is None
    assert len(a.assign(torch.empty((0, 4)), torch.empty((0, 4))).gt_inds) == 0      # :151-162

