"""CPU: the C-ABI shared library loads and exports every symbol include/aidet_b200.h declares
(no compute calls without a GPU), and the Python mirror keeps the reference's surface."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

import aidet_b200
from aidet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "aidet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aidet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = _declared_symbols()
    assert len(names) >= 12
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "header declares %s but the library does not export it" % n
    assert sorted(_lib.EXPORTS) == names, "aidet_b200/_lib.py and include/aidet_b200.h disagree"


def test_library_loads_and_reports_errors_without_gpu():
    lib = _lib.lib()
    assert lib.aidet_version() >= 100
    assert lib.aidet_riou_workspace_bytes(2000, 2000, 5) >= 2 * 2000 * 32
    assert lib.aidet_riou_workspace_bytes(10, 10, 8) >= 2 * 10 * 64
    # argument validation happens before any CUDA call, so it is observable on a CPU-only box
    rc = lib.aidet_riou_matrix_f32(None, 4, None, 4, 7, 0, None, 4, None, 0, 0, None)
    assert rc == -1 and b"fmt" in lib.aidet_last_error()
    rc = lib.aidet_nms_batched_f32(None, 5, None, None, 4, None, 3, 2, 0, 0, None, None, None, 0, 0, None)
    assert rc == -1 and b"n_thr" in lib.aidet_last_error()
    rc = lib.aidet_rroi_align_fwd_f32(None, None, None, None, 9, 1, 1, None, 5, None, 0, 7, 7, 2, 0, None, 0, None)
    assert rc == -1 and b"n_levels" in lib.aidet_last_error()
    rc = lib.aidet_max_iou_assign_f32(None, 0, None, 4, 5, None, 0, 0.0, 0, 0.5, 0.0, 0.5, 0.0, 1, None, None, None, None,
                                      None, 0, 0, None)
    assert rc == -1 and b"at least one gt" in lib.aidet_last_error()
    rc = lib.aidet_max_iou_assign_f32(None, 3, None, 4, 6, None, 0, 0.0, 0, 0.5, 0.0, 0.5, 0.0, 1, None, None, None, None,
                                      None, 0, 0, None)
    assert rc == -1 and b"fmt" in lib.aidet_last_error()
    rc = lib.aidet_assign_wrt_overlaps_f32(None, 3, 4, 2, 0.5, 0.0, 0.5, 0.0, 1, None, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.aidet_last_error()
    rc = lib.aidet_riou_aligned_grad_f32(None, None, 4, 4, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"fmt must be 5" in lib.aidet_last_error()
    # workspace sizes grow with every operand and cover the records (32 B per theta-OBB)
    w0 = lib.aidet_assign_workspace_bytes(100, 2000, 0, 5)
    assert w0 >= (100 + 2000) * 32 + 2000 * 16 and lib.aidet_assign_workspace_bytes(100, 2000, 7, 5) > w0
    assert lib.aidet_assign_workspace_bytes(100, 2000, 0, 0) < w0          # matrix form: no records
    assert lib.aidet_launch_count() == 0


def test_python_surface_matches_reference_names():
    from aidet_b200 import core, ops
    from aidet_b200.ops.nms import nms_wrapper
    for name in ("nms", "soft_nms", "thetaobb_nms", "pointobb_nms", "batched_rnms"):
        assert callable(getattr(nms_wrapper, name))
    assert list(inspect.signature(nms_wrapper.nms).parameters) == ["dets", "iou_thr", "device_id"]
    assert list(inspect.signature(nms_wrapper.thetaobb_nms).parameters) == ["dets", "iou_thr", "device_id"]
    sig = inspect.signature(ops.RoIAlign.__init__).parameters
    assert list(sig) == ["self", "out_size", "spatial_scale", "sample_num", "use_torchvision", "aligned"]
    assert sig["aligned"].default is False and sig["sample_num"].default == 0
    m = ops.RoIAlign(7, 1 / 16, sample_num=2)
    assert m.out_size == (7, 7) and m.spatial_scale == 1 / 16
    assert repr(m) == "RoIAlign(out_size=(7, 7), spatial_scale=0.0625, sample_num=2, use_torchvision=False, aligned=False)"
    # name-based lookups used by the reference (single_level.py:47-49, rbbox_nms.py:26-27)
    assert getattr(ops, "RoIAlign") is ops.RoIAlign and getattr(ops, "RoIAlignRotated")
    assert list(inspect.signature(core.rbbox_overlaps).parameters) == ["rbboxes1", "rbboxes2", "mode", "is_aligned"]
    assert list(inspect.signature(core.multiclass_thetaobb_nms).parameters)[:5] == [
        "multi_rbboxes", "multi_scores", "score_thr", "polygon_nms_iou_thr", "max_num"]
    # max_iou_assigner.py:37-44, iou_loss.py:131
    assert list(inspect.signature(core.MaxIoUAssigner.__init__).parameters) == [
        "self", "pos_iou_thr", "neg_iou_thr", "min_pos_iou", "gt_max_assign_all", "ignore_iof_thr", "ignore_wrt_candidates",
        "gpu_assign_thr"]
    assert list(inspect.signature(core.MaxIoUAssigner.assign).parameters) == [
        "self", "bboxes", "gt_bboxes", "gt_bboxes_ignore", "gt_labels"]
    from aidet_b200 import models
    assert list(inspect.signature(models.RotatedIoULoss.__init__).parameters) == ["self", "eps", "reduction", "loss_weight"]
    # core/rbbox/transforms.py:191,398,405 and rbbox_target.py:8-17
    assert list(inspect.signature(core.thetaobb_flip).parameters) == ["thetaobbs", "img_shape"]
    assert list(inspect.signature(core.thetaobb_mapping_back).parameters) == ["thetaobbs", "img_shape", "scale_factor", "flip"]
    assert list(inspect.signature(core.rbbox_target).parameters) == [
        "pos_proposals_list", "neg_proposals_list", "pos_assigned_gt_inds_list", "gt_rbboxes_list", "gt_labels_list",
        "rbbox_test_cfg", "target_means", "target_stds", "out_dim_reg", "concat"]


def test_no_cpu_fallback():
    """CPU tensors must raise, never silently compute (reference: roi_align.py:41-42)."""
    from aidet_b200 import core, ops
    with pytest.raises(NotImplementedError):
        ops.roi_align(torch.randn(1, 4, 8, 8), torch.zeros(1, 5), 3, 0.25, 2, True)
    with pytest.raises(NotImplementedError):
        ops.roi_align_rotated(torch.randn(1, 4, 8, 8), torch.zeros(1, 6), 3, 0.25, 2, True)
    with pytest.raises(NotImplementedError):
        core.rbbox_overlaps(torch.rand(3, 5), torch.rand(3, 5))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ops.nms(np.zeros((3, 5), np.float32), 0.5)
    with pytest.raises(TypeError):
        ops.nms([[0, 0, 1, 1, 0.5]], 0.5)
    # empty inputs keep the reference's shapes without touching the device (nms_wrapper.py:50-51, geometry.py:54-55)
    kept, inds = ops.thetaobb_nms(torch.zeros((0, 6)), 0.5)
    assert kept.shape == (0, 6) and inds.dtype == torch.long and inds.shape == (0,)
    assert tuple(core.rbbox_overlaps(torch.zeros((0, 5)), torch.zeros((3, 5))).shape) == (0, 3)
    assert tuple(core.rbbox_overlaps(torch.zeros((0, 5)), torch.zeros((0, 5)), is_aligned=True).shape) == (0, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.dirname(aidet_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_install_as_mmdet_namespace():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "mmdet" or k.startswith("mmdet.")}
    try:
        for k in list(saved):
            del sys.modules[k]
        aidet_b200.install_as_mmdet()
        from mmdet.ops import RoIAlign, nms  # noqa: F401
        from mmdet.ops.nms import nms_wrapper
        from mmdet.core import rbbox_overlaps  # noqa: F401
        assert getattr(nms_wrapper, "thetaobb_nms") is aidet_b200.ops.thetaobb_nms
    finally:
        for k in [k for k in sys.modules if k == "mmdet" or k.startswith("mmdet.")]:
            del sys.modules[k]
        sys.modules.update(saved)
