"""GPU: the fused max-IoU assignment (aidet_max_iou_assign_f32 / aidet_assign_wrt_overlaps_f32 through the
MaxIoUAssigner mirror) against (1) outputs of the reference's own MaxIoUAssigner (tests/golden/golden_assign_v1.npz)
and (2) the numpy restatement of it (oracle/oracle.py) run on the device's own overlap matrix -- bit-exact -- and on
the float64 oracle overlaps (boxes whose decision lies within 1e-6 of a threshold or a tie excluded and counted)."""
import os

import numpy as np
import pytest
import torch

from aidet_b200 import synth
from aidet_b200.core import MaxIoUAssigner, bbox_overlaps, rbbox_overlaps
from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_assign_v1.npz"))
N_CFG = GOLD["configs"].shape[0]


def golden_cfg(ci):
    pos, nlo, nhi, mp, allg, ign, wrt, is_pair = GOLD["configs"][ci]
    neg = (float(nlo), float(nhi)) if is_pair else float(nhi)
    return float(pos), neg, float(mp), bool(allg), float(ign), bool(wrt)


def test_reference_doctest(cuda):
    # max_iou_assigner.py:78-84
    r = MaxIoUAssigner(0.5, 0.5).assign(torch.Tensor([[0, 0, 10, 10], [10, 10, 20, 20]]).to(cuda),
                                        torch.Tensor([[0, 0, 10, 9]]).to(cuda))
    assert r.gt_inds.tolist() == [1, 0] and r.num_gts == 1 and r.labels is None


def test_reference_assigner_tests(cuda):
    """The reference's own tests/test_assigner.py:17-65 (golden gt_inds) through the device path."""
    bboxes = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]]).to(cuda)
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]]).to(cuda)
    r = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5).assign(bboxes, gts, gt_labels=torch.LongTensor([2, 3]).to(cuda))
    assert len(r.gt_inds) == 4 and len(r.labels) == 4
    assert torch.all(r.gt_inds.cpu() == torch.LongTensor([1, 0, 2, 0])) and r.labels.tolist() == [2, 0, 3, 0]
    a = MaxIoUAssigner(pos_iou_thr=0.5, neg_iou_thr=0.5, ignore_iof_thr=0.5, ignore_wrt_candidates=False)
    r = a.assign(bboxes, gts, gt_bboxes_ignore=torch.Tensor([[30, 30, 40, 40]]).to(cuda))
    assert torch.all(r.gt_inds.cpu() == torch.LongTensor([1, 0, 2, -1]))


@pytest.mark.parametrize("ci", range(N_CFG))
def test_golden_overlap_matrices(cuda, ci):
    """assign_wrt_overlaps on the reference's matrices (exact ties, -1 columns, an all-zero gt row): bit-exact."""
    pos, neg, mp, allg, _, _ = golden_cfg(ci)
    a = MaxIoUAssigner(pos, neg, mp, allg)
    r = a.assign_wrt_overlaps(torch.from_numpy(GOLD["c%d_mat" % ci]).to(cuda),
                              torch.from_numpy(GOLD["c%d_gt_labels" % ci][:17]).to(cuda))
    assert np.array_equal(r.gt_inds.cpu().numpy(), GOLD["c%d_mat_gt_inds" % ci])
    assert np.array_equal(r.max_overlaps.cpu().numpy(), GOLD["c%d_mat_max_overlaps" % ci])
    assert np.array_equal(r.labels.cpu().numpy(), GOLD["c%d_mat_labels" % ci])


@pytest.mark.parametrize("ci", range(N_CFG))
def test_golden_hbb_boxes(cuda, ci):
    """assign on axis-aligned boxes (+1 convention) with ignore boxes and labels vs the reference's own run."""
    pos, neg, mp, allg, ign, wrt = golden_cfg(ci)
    a = MaxIoUAssigner(pos, neg, mp, allg, ign, wrt)
    r = a.assign(torch.from_numpy(GOLD["c%d_boxes" % ci]).to(cuda), torch.from_numpy(GOLD["c%d_gts" % ci]).to(cuda),
                 torch.from_numpy(GOLD["c%d_ign" % ci]).to(cuda), torch.from_numpy(GOLD["c%d_gt_labels" % ci]).to(cuda))
    assert np.abs(r.max_overlaps.cpu().numpy() - GOLD["c%d_max_overlaps" % ci]).max() < 1e-6
    assert np.array_equal(r.gt_inds.cpu().numpy(), GOLD["c%d_gt_inds" % ci])
    assert np.array_equal(r.labels.cpu().numpy(), GOLD["c%d_labels" % ci])


CASES = [  # n_boxes, n_gts, n_ignore, (pos, neg, min_pos, assign_all, ign_thr, wrt_candidates)
    (2000, 37, 0, (0.5, 0.5, 0.5, True, -1, True)),            # rcnn: 2000 proposals (configs/*: assigner of rcnn)
    (2000, 37, 6, (0.5, 0.5, 0.0, True, 0.5, True)),
    (1, 1, 0, (0.5, 0.5, 0.0, True, -1, True)),
    (300, 700, 9, (0.7, (0.1, 0.3), 0.3, False, 0.4, False)),  # more truths than boxes, several row tiles
    (50000, 130, 4, (0.7, 0.3, 0.3, True, 0.5, True)),         # rpn-sized candidate set
]


@pytest.mark.parametrize("fmt", [5, 8])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_rotated_assign_vs_oracle(cuda, case, fmt):
    nb, ng, ni, (pos, neg, mp, allg, ign, wrt) = CASES[case]
    boxes, gts, ignb, labels = synth.assign_case(nb, ng, seed=50 + case, n_ignore=ni)
    if fmt == 8:
        boxes, gts, ignb = synth.thetaobb2pointobb(boxes), synth.thetaobb2pointobb(gts), synth.thetaobb2pointobb(ignb)
    a = MaxIoUAssigner(pos, neg, mp, allg, ign, wrt)
    r = a.assign(boxes.to(cuda), gts.to(cuda), ignb.to(cuda) if ni else None, labels.to(cuda))
    gi, mo, lb = r.gt_inds.cpu().numpy(), r.max_overlaps.cpu().numpy(), r.labels.cpu().numpy()
    assert r.num_gts == ng and gi.shape == (nb,) and gi.dtype == np.int64

    # (1) the reference steps on the DEVICE's own overlap matrices: bit-exact (ties included)
    ov = rbbox_overlaps(gts.to(cuda), boxes.to(cuda)).cpu().numpy()
    if ni:
        if wrt:
            ign_max = rbbox_overlaps(boxes.to(cuda), ignb.to(cuda), mode='iof').cpu().numpy().max(1)
        else:
            ign_max = rbbox_overlaps(ignb.to(cuda), boxes.to(cuda), mode='iof').cpu().numpy().max(0)
        ov[:, ign_max > ign] = -1
    gi1, mo1, lb1 = O.max_iou_assign_wrt_overlaps(ov, pos, neg, mp, allg, labels.numpy())
    if fmt == 5:
        assert np.array_equal(mo, mo1)
        assert np.array_equal(gi, gi1)
        assert np.array_equal(lb, lb1)
    else:
        # point-OBB: the fused kernel and the matrix kernel are different instantiations of the pair arithmetic and
        # may differ in the last bit, so a decision within 1e-6 of a threshold may flip; everything else is equal
        assert np.abs(mo - mo1).max() <= 1e-6
        bad = (gi != gi1) | (lb != lb1)
        assert bad.sum() <= 2 and (mo[bad] != mo1[bad]).all()

    # (2) float64 oracle overlaps: identical except where a decision sits within 1e-6 of a threshold / tie
    gi2, mo2, _ = O.max_iou_assign(boxes.numpy(), gts.numpy(), pos, neg, mp, allg, ign, wrt,
                                   ignb.numpy() if ni else None, labels.numpy())
    ok = mo2 >= 0
    assert np.abs(mo - mo2)[ok & (mo >= 0)].max() <= 1e-5
    diff = gi != gi2
    assert diff.mean() <= 2e-3, "%d of %d boxes differ from the float64 oracle" % (diff.sum(), nb)
    assert (gi > 0).sum() > 0 or nb == 1


def test_assign_wrt_overlaps_equals_fused(cuda):
    boxes, gts, _, labels = synth.assign_case(5000, 90, seed=9)
    a = MaxIoUAssigner(0.5, 0.4, 0.3, True)
    r1 = a.assign(boxes.to(cuda), gts.to(cuda), None, labels.to(cuda))
    r2 = a.assign_wrt_overlaps(rbbox_overlaps(gts.to(cuda), boxes.to(cuda)), labels.to(cuda))
    assert torch.equal(r1.gt_inds, r2.gt_inds) and torch.equal(r1.max_overlaps, r2.max_overlaps)
    assert torch.equal(r1.labels, r2.labels)
    # HBB form of the same consistency (bbox_overlaps, +1)
    hb = torch.cat([boxes[:, :2] - boxes[:, 2:4] / 2, boxes[:, :2] + boxes[:, 2:4] / 2], 1)
    hg = torch.cat([gts[:, :2] - gts[:, 2:4] / 2, gts[:, :2] + gts[:, 2:4] / 2], 1)
    r1 = a.assign(hb.to(cuda), hg.to(cuda))
    r2 = a.assign_wrt_overlaps(bbox_overlaps(hg.to(cuda), hb.to(cuda)))
    assert torch.equal(r1.gt_inds, r2.gt_inds) and torch.equal(r1.max_overlaps, r2.max_overlaps)


def test_scores_column_is_dropped_and_errors(cuda):
    boxes, gts, _, _ = synth.assign_case(100, 5, seed=3)
    a = MaxIoUAssigner(0.5, 0.5)
    with_score = torch.cat([boxes, torch.rand(100, 1)], 1)          # proposals carry a score column (:101)
    assert torch.equal(a.assign(with_score.to(cuda), gts.to(cuda)).gt_inds, a.assign(boxes.to(cuda), gts.to(cuda)).gt_inds)
    with pytest.raises(AssertionError):
        a.assign(boxes.to(cuda), torch.rand(3, 6).to(cuda))



@pytest.mark.gpu
def test_negative_zero_entries_in_a_caller_matrix(cuda):
    """-0.0 is a legal overlap value in assign_wrt_overlaps: it must rank as 0, not above every positive entry (its
    bits are 0x80000000), exactly as the numpy restatement of max_iou_assigner.py:122-195 ranks it."""
    rng = np.random.default_rng(3)
    ov = (np.round(rng.uniform(0, 1, (6, 200)) * 10) / 10).astype(np.float32)
    ov[:, ::7] = -0.0
    ov[2] = np.where(np.arange(200) % 2 == 0, -0.0, 0.0).astype(np.float32)          # a truth whose best overlap is zero
    lab = np.arange(1, 7)
    a = MaxIoUAssigner(0.5, 0.4, 0.0, True)
    r = a.assign_wrt_overlaps(torch.from_numpy(ov).to(cuda), torch.from_numpy(lab).to(cuda))
    gi, mo, lb = O.max_iou_assign_wrt_overlaps(ov, 0.5, 0.4, 0.0, True, lab)
    assert np.array_equal(r.gt_inds.cpu().numpy(), gi) and np.array_equal(r.labels.cpu().numpy(), lb)
    assert np.array_equal(r.max_overlaps.cpu().numpy(), np.abs(mo))                   # -0.0 == 0.0 by value
