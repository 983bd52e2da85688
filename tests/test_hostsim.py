"""CPU: the FP32 pair arithmetic of aidet_b200/csrc/geom.cuh, compiled for the host (tests/hostsim/),
against the float64 oracle.  This checks the algorithm the kernels run without a GPU; it is a
test harness, not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from aidet_b200 import synth
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "geom_sim.cpp")
SO = os.path.join(HERE, "hostsim", "libgeom_sim.so")
HDR = os.path.join(HERE, "..", "aidet_b200", "csrc", "geom.cuh")


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=fast", "-mfma", "-o", SO, SRC])
    lib = C.CDLL(SO)

    def run(a, b, mode=0):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.empty((len(a), len(b)), np.float32)
        lib.sim_riou_matrix(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b),
                            a.shape[1], mode, out.ctypes.data_as(C.c_void_p))
        return out.astype(np.float64)
    return run


@pytest.mark.parametrize("dense", [False, True])
def test_rect_iou_and_iof(sim, dense):
    a, _ = synth.dota_boxes(700, seed=31, dense=dense)
    b, _ = synth.dota_boxes(700, seed=32, dense=dense)
    a, b = a.numpy(), b.numpy()
    assert np.abs(sim(a, b) - O.riou_matrix(a, b)).max() < 2e-6
    assert np.abs(sim(a, b, 1) - O.riou_matrix(a, b, mode="iof")).max() < 1e-5


def test_quad_iou_and_iof(sim):
    a, _ = synth.dota_boxes(500, side=600, seed=33)
    b, _ = synth.dota_boxes(500, side=600, seed=34)
    a8, ca = synth.free_quads(a, 0.1, seed=1)
    b8, cb = synth.free_quads(b, 0.1, seed=2)
    a8, b8 = a8[ca].numpy(), b8[cb].numpy()
    assert len(a8) > 400 and len(b8) > 400
    assert np.abs(sim(a8, b8) - O.riou_matrix(a8, b8)).max() < 5e-6
    assert np.abs(sim(a8, b8, 1) - O.riou_matrix(a8, b8, mode="iof")).max() < 3e-5
    b8cw = np.ascontiguousarray(b8.reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8))
    assert np.abs(sim(a8, b8cw) - sim(a8, b8)).max() < 5e-6


def test_nonconvex_quads_follow_the_fan_algorithm(sim):
    """Simple non-convex quads: the signed-triangle decomposition (DOTA_devkit polyiou lineage, oracle
    ALGO_FAN) is the definition; Sutherland-Hodgman needs a convex clip polygon and does not apply."""
    dart = np.array([[0, 0, 10, 0, 3, 3, 0, 10]], np.float32)          # reflex vertex at (3,3)
    sq = np.array([[1, 1, 9, 1, 9, 9, 1, 9], [0, 0, 4, 0, 4, 4, 0, 4], [5, 5, 12, 5, 12, 12, 5, 12]], np.float32)
    for x, y in ((dart, sq), (sq, dart)):
        assert np.abs(sim(x, y) - O.riou_matrix(x, y, algo=O.ALGO_FAN)).max() < 2e-6


def test_degenerate_boxes_are_finite(sim):
    """zero-size, sliver and identical boxes: no NaN/inf, IoU in [0,1]."""
    sp = np.array([[0, 0, 2, 2, 0], [0, 0, 2, 2, np.pi / 4], [0, 0, 0, 0, 0], [5, 5, 1e-3, 100, 1.0],
                   [0, 0, 2, 0, 0.3], [1e4, 1e4, 50, 20, -0.7], [1e4, 1e4, 50, 20, -0.7], [0, 0, 2, 2, np.pi / 2]],
                  np.float32)
    m = sim(sp, sp)
    assert np.isfinite(m).all() and (m >= 0).all() and (m <= 1).all()
    ref = O.riou_matrix(sp, sp)
    keep = [0, 1, 5, 6, 7]
    assert np.abs(m[np.ix_(keep, keep)] - ref[np.ix_(keep, keep)]).max() < 2e-6
    assert m[2].max() == 0 and m[:, 2].max() == 0


@pytest.mark.parametrize("side,dense", [(600, False), (600, True), (16384, True), (16384, False)])
def test_rectangles_given_as_corners_take_the_parallelogram_path(sim, side, dense):
    """8-point boxes that are rectangles (thetaobb2pointobb output, float32-rounded corners, coordinates up to 16384 as in
    config C4): the affine-square integral of geom.cuh: para_inter must agree with the float64 oracle on the SAME
    rounded corners within the IoU tolerance, like the general fan does."""
    a, _ = synth.dota_boxes(400, side=side, seed=41, dense=dense)
    b, _ = synth.dota_boxes(400, side=side, seed=42, dense=dense)
    a8, b8 = synth.thetaobb2pointobb(a).float().numpy(), synth.thetaobb2pointobb(b).float().numpy()
    ref = O.riou_matrix(a8, b8)
    got = sim(a8, b8)
    assert np.abs(got - ref).max() < (5e-6 if side < 10000 else 1e-5), np.abs(got - ref).max()
    assert np.abs(sim(a8, b8, 1) - O.riou_matrix(a8, b8, mode="iof")).max() < 3e-5
    assert (ref > 0).sum() > (100000 if dense else (1000 if side < 10000 else -1))
    # clockwise corner order and a rotated starting corner describe the same box
    b8cw = np.ascontiguousarray(b8.reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8))
    b8rot = np.ascontiguousarray(np.roll(b8.reshape(-1, 4, 2), 1, axis=1).reshape(-1, 8))
    assert np.abs(sim(a8, b8cw) - got).max() < 5e-6 and np.abs(sim(a8, b8rot) - got).max() < 5e-6


def test_parallelograms_and_mixed_sets(sim):
    """Sheared boxes (parallelograms that are not rectangles) take the same path; sets that mix parallelograms with free
    quads dispatch per pair and every combination agrees with the oracle; identical boxes give 1."""
    a, _ = synth.dota_boxes(300, side=500, seed=43)
    b, _ = synth.dota_boxes(300, side=500, seed=44)
    pa, pb = synth.thetaobb2pointobb(a).double().numpy().reshape(-1, 4, 2), synth.thetaobb2pointobb(b).double().numpy().reshape(-1, 4, 2)
    rng = np.random.default_rng(0)
    def shear(p):
        c = p.mean(1, keepdims=True)
        k = rng.uniform(-0.6, 0.6, (p.shape[0], 1, 1))
        q = p - c
        q = np.concatenate([q[..., :1] + k * q[..., 1:], q[..., 1:]], -1)        # x += k y: still a parallelogram
        return (q + c).reshape(-1, 8).astype(np.float32)
    sa, sb = shear(pa), shear(pb)
    assert np.abs(sim(sa, sb) - O.riou_matrix(sa, sb)).max() < 5e-6
    fa, ca = synth.free_quads(a, 0.1, seed=3)
    mixed = sa.copy()
    mixed[::2] = fa.numpy()[::2]
    okm = np.ones(len(mixed), bool); okm[::2] = ca.numpy()[::2]
    assert np.abs(sim(mixed[okm], sb) - O.riou_matrix(mixed[okm], sb)).max() < 5e-6
    assert np.abs(sim(sb, mixed[okm]) - O.riou_matrix(sb, mixed[okm])).max() < 5e-6
    d = np.diag(sim(sa, sa))
    assert np.abs(d - 1).max() < 2e-6
    # degenerate parallelograms: zero area, slivers
    deg = np.array([[0, 0, 4, 0, 4, 0, 0, 0], [0, 0, 4, 0, 4, 1e-3, 0, 1e-3], [1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 4, 0, 4, 4, 0, 4]], np.float32)
    m = sim(deg, deg)
    assert np.isfinite(m).all() and (m >= 0).all() and (m <= 1).all() and abs(m[3, 3] - 1) < 1e-6 and m[0].max() == 0
