"""CPU: the OBB codecs, flips, TTA mappings and rbbox_target of aidet_b200.core.rbbox against OUTPUTS OF THE REFERENCE'S
OWN FUNCTIONS (tests/golden/golden_rbbox_v1.npz, made by tests/golden/make_golden_rbbox.py from
/root/reference/mmdet/core/rbbox/{transforms,rbbox_target}.py) -- pinned, not restated."""
import os

import numpy as np
import pytest
import torch

from aidet_b200 import core

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rbbox_v1.npz"))
IMG = tuple(int(v) for v in G["img_shape"])


def close(a, b, tol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()) if a.size else True


def test_list_converters_one_box_at_a_time():
    # transforms.py:30-163: list in -> list out
    for i in range(G["theta_in"].shape[0]):
        th, pt = G["theta_in"][i].tolist(), G["theta2point"][i].tolist()
        got = core.thetaobb2pointobb(th)
        assert isinstance(got, list) and close(got, G["theta2point"][i], 0)                 # same cv2 call: bit-equal
        got = core.pointobb2thetaobb(pt)
        assert isinstance(got, list) and close(got, G["point2theta"][i], 0)
        assert core.pointobb2bbox(pt) == G["point2bbox"][i].tolist()
        assert core.pointobb_extreme_sort(pt) == G["extreme_sort"][i].tolist()
        assert core.pointobb_best_point_sort(pt) == G["best_point_sort"][i].tolist()
        assert close(core.thetaobb2hobb(th, core.pointobb_best_point_sort), G["theta2hobb_best"][i], 0)
        assert close(core.thetaobb2hobb(th, core.pointobb_extreme_sort), G["theta2hobb_extreme"][i], 0)
        assert core.hobb2pointobb(G["theta2hobb_best"][i].tolist()) == G["hobb2point"][i].astype(int).tolist()


def test_batched_converters_equal_the_per_box_goldens():
    assert close(core.thetaobb2pointobb(G["theta_in"]), G["theta2point"], 0)
    assert close(core.pointobb2thetaobb(G["theta2point"]), G["point2theta"], 0)
    assert close(core.pointobb_extreme_sort(G["theta2point"]), G["extreme_sort"], 0)
    assert close(core.pointobb_best_point_sort(G["theta2point"]), G["best_point_sort"], 0)
    assert close(core.thetaobb2hobb(G["theta_in"], core.pointobb_extreme_sort), G["theta2hobb_extreme"], 0)
    # a user-supplied sort callable goes through the per-box path
    assert close(core.thetaobb2hobb(G["theta_in"], lambda p: core.pointobb_best_point_sort(p)), G["theta2hobb_best"], 0)
    # ties: equally high top points -> the left one (transforms.py:105-109); equal distances -> the smaller shift
    assert close(core.pointobb_extreme_sort(G["tie_in"]), G["tie_extreme_sort"], 0)
    assert close(core.pointobb_best_point_sort(G["tie_in"]), G["tie_best_point_sort"], 0)
    # the tensor form of thetaobb2pointobb (closed form, used on the device) agrees with cv2's float32 to rounding
    t = core.thetaobb2pointobb(torch.from_numpy(G["theta_in"]))
    assert close(t.numpy(), G["theta2point"], 1e-6)


def test_flips():
    th, pt, hb = G["theta_in"], G["theta2point"], G["hobb_in"]
    assert close(core.thetaobb_flip(th, IMG), G["thetaobb_flip"], 0)
    assert close(core.thetaobb_flip(th[0], IMG), G["thetaobb_flip_1d"], 0)
    assert close(core.pointobb_flip(pt, IMG), G["pointobb_flip"], 0)
    assert close(core.pointobb_flip(pt[5], IMG), G["pointobb_flip_1d"], 0)
    assert close(core.hobb_flip(hb, IMG), G["hobb_flip"], 0)
    assert close(core.hobb_flip(hb[3], IMG), G["hobb_flip_1d"], 0)                           # (1, 5), as the reference
    src = th.copy()
    core.thetaobb_flip(src, IMG)
    assert np.array_equal(src, th)                                                            # not in place (:199 copies)
    # tensors take the same route (decoded detections are flipped back on the device)
    tt = core.thetaobb_flip(torch.from_numpy(th), IMG)
    assert isinstance(tt, torch.Tensor) and close(tt.numpy(), G["thetaobb_flip"], 0)
    tp = core.pointobb_flip(torch.from_numpy(pt).float(), IMG)
    assert isinstance(tp, torch.Tensor) and tp.dtype == torch.float32 and close(tp.numpy(), G["pointobb_flip"], 1e-6)


@pytest.mark.parametrize("flip", [False, True])
def test_tta_mappings(flip):
    tag = "_flip" if flip else ""
    th, pt, hb = G["theta_in"], G["theta2point"], G["hobb_in"]
    assert close(core.thetaobb_mapping(th, IMG, 1.5, flip), G["thetaobb_mapping" + tag], 0)
    assert close(core.thetaobb_mapping_back(th, IMG, 1.5, flip), G["thetaobb_mapping_back" + tag], 1e-15)
    assert close(core.pointobb_mapping(pt, IMG, 0.75, flip), G["pointobb_mapping" + tag], 0)
    assert close(core.pointobb_mapping_back(pt, IMG, 0.75, flip), G["pointobb_mapping_back" + tag], 1e-15)
    assert close(core.hobb_mapping(hb, IMG, 1.25, flip), G["hobb_mapping" + tag], 0)
    assert close(core.hobb_mapping_back(hb, IMG, 1.25, flip), G["hobb_mapping_back" + tag], 1e-15)


def test_rescale_and_delta_codecs():
    m = torch.from_numpy(G["thetaobb_rescale_in"]).clone()
    assert close(core.thetaobb_rescale(m.clone(), 2.5).numpy(), G["thetaobb_rescale"], 0)
    assert close(core.thetaobb_rescale(m.clone(), 2.5, reverse_flag=True).numpy(), G["thetaobb_rescale_rev"], 0)
    prop = torch.from_numpy(G["prop"])
    m5, s5 = [0.0] * 5, G["stds5"].tolist()
    m8, s8 = [0.0] * 8, G["stds8"].tolist()
    tol = 2e-6                                                                                # float32, different op order
    d5 = core.thetaobb2delta(prop, torch.from_numpy(G["theta_in"]).float(), m5, s5)
    assert close(d5.numpy(), G["thetaobb2delta"], tol)
    assert close(core.delta2thetaobb(prop, torch.from_numpy(G["thetaobb2delta"]), m5, s5).numpy(), G["delta2thetaobb"], tol)
    assert close(core.delta2thetaobb(prop, torch.from_numpy(G["delta5_multi"]), m5, s5).numpy(), G["delta2thetaobb_multi"], tol)
    d8 = core.pointobb2delta(prop, torch.from_numpy(G["theta2point"]).float(), m8, s8)
    assert close(d8.numpy(), G["pointobb2delta"], tol)
    assert close(core.delta2pointobb(prop, torch.from_numpy(G["pointobb2delta"]), m8, s8).numpy(), G["delta2pointobb"], tol)
    assert close(core.delta2pointobb(prop, torch.from_numpy(G["delta8_multi"]), m8, s8).numpy(), G["delta2pointobb_multi"], tol)
    dh = core.hobb2delta(prop, torch.from_numpy(G["hobb_in"]).float(), m5, s5)
    assert close(dh.numpy(), G["hobb2delta"], tol)
    assert close(core.delta2hobb(prop, torch.from_numpy(G["hobb2delta"]), m5, s5).numpy(), G["delta2hobb"], tol)
    assert close(core.delta2hobb(prop, torch.from_numpy(G["deltah_multi"]), m5, s5).numpy(), G["delta2hobb_multi"], tol)


class _Cfg:
    def __init__(self, encode):
        self.encode = encode


@pytest.mark.parametrize("encode,dim,src", [("thetaobb", 5, "theta_in"), ("pointobb", 8, "theta2point"), ("hobb", 5, "hobb_in")])
def test_rbbox_target(encode, dim, src):
    # rbbox_target.py:8-88: two images, the second without negatives
    prop = torch.from_numpy(G["prop"])
    gts = torch.from_numpy(G[src]).float()
    pos, neg = [prop[:9], prop[40:45]], [prop[20:31], prop[:0]]
    inds = [torch.from_numpy(G["target_inds1"]), torch.from_numpy(G["target_inds2"])]
    labels = [torch.from_numpy(G["target_labels1"]), torch.from_numpy(G["target_labels2"])]
    means = [0.0] * dim
    stds = (G["stds8"] if dim == 8 else G["stds5"]).tolist()
    for cfg in (_Cfg(encode), {"encode": encode}):
        got = core.rbbox_target(pos, neg, inds, [gts[:8], gts[30:35]], labels, cfg, means, stds, out_dim_reg=dim)
        assert got[0].dtype == torch.long and np.array_equal(got[0].numpy(), G["target_%s_labels" % encode])
        assert np.array_equal(got[1].numpy(), G["target_%s_label_weights" % encode])
        assert close(got[2].numpy(), G["target_%s_targets" % encode], 2e-6)
        assert np.array_equal(got[3].numpy(), G["target_%s_weights" % encode])
    split = core.rbbox_target(pos, neg, inds, [gts[:8], gts[30:35]], labels, _Cfg(encode), means, stds, out_dim_reg=dim,
                              concat=False)
    assert [t.shape[0] for t in split[0]] == G["target_%s_split_sizes" % encode].tolist()
    # no positives at all: everything zero but the label weights of the negatives
    e = core.rbbox_target_single(prop[:0], prop[:4], torch.zeros(0, dtype=torch.long), gts[:8], labels[0], _Cfg(encode),
                                 means, stds, out_dim_reg=dim)
    assert e[0].tolist() == [0] * 4 and e[1].tolist() == [1.0] * 4 and float(e[2].abs().sum() + e[3].abs().sum()) == 0.0


def test_rbbox2roi():
    rr = core.rbbox2roi([torch.zeros(0, 6), torch.tensor([[1., 2., 3., 4., .5, .9]]), torch.tensor([[5., 6., 7., 8., -.1]])])
    assert rr.shape == (2, 6) and rr[0].tolist() == [1, 1, 2, 3, 4, .5]
    assert [round(v, 4) for v in rr[1].tolist()] == [2, 5, 6, 7, 8, -.1]
    assert core.rbbox2roi([torch.zeros(0, 5)]).shape == (0, 6)
