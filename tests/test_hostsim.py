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
