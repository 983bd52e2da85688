"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module.  The product package `aidet_b200` never does; it fails loudly
when its CUDA library is missing instead of falling back to anything in here.

Parity status: rotated IoU / polygon NMS are "parity unpinned" (the arithmetic lives
in the un-vendored third-party `wwtool`, see oracle_geom.c header); HBB NMS is pinned
to the reference's own nms_cpu.cpp (oracle/_ref) and RoIAlign to torchvision CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

ALGO_SH, ALGO_FAN = 0, 1
MODE_IOU, MODE_IOF = 0, 1
ROI_V1, ROI_V2, ROI_V2_ALIGNED = 0, 1, 2


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle_geom.c", "oracle_roi.c")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_riou_pair.restype = C.c_double
        _lib.oracle_riou_pair_d.restype = C.c_double
        _lib.oracle_nms.restype = C.c_int
        _lib.oracle_nms_verify.restype = C.c_int
    return _lib


def _f32(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def thetaobb2pointobb(boxes):
    b = _f32(boxes, 5)
    out = np.empty((b.shape[0], 8), np.float64)
    lib().oracle_thetaobb2pointobb(_p(b), C.c_int(b.shape[0]), _p(out))
    return out


def riou_matrix(a, b, mode="iou", algo=ALGO_SH):
    fmt = np.asarray(a).shape[-1] if np.asarray(a).size else np.asarray(b).shape[-1]
    a, b = _f32(a, fmt), _f32(b, fmt)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    lib().oracle_riou_matrix(_p(a), C.c_int(a.shape[0]), _p(b), C.c_int(b.shape[0]), C.c_int(fmt),
                             C.c_int(MODE_IOF if mode == "iof" else MODE_IOU), C.c_int(algo), _p(out))
    return out


def riou_aligned(a, b, mode="iou", algo=ALGO_SH):
    fmt = np.asarray(a).shape[-1]
    a, b = _f32(a, fmt), _f32(b, fmt)
    assert a.shape == b.shape
    out = np.empty((a.shape[0],), np.float64)
    lib().oracle_riou_aligned(_p(a), _p(b), C.c_int(a.shape[0]), C.c_int(fmt),
                              C.c_int(MODE_IOF if mode == "iof" else MODE_IOU), C.c_int(algo), _p(out))
    return out


def riou_aligned_grad_fd(a, b, mode="iou", step=1e-6):
    """theta-OBB (n,5) or point-OBB (n,8) pairs in float64 -> (overlap (n,), central-difference gradient (n, 2*fmt)
    w.r.t. a's then b's parameters).  The checker for the analytic gradient of the rotated IoU loss."""
    fmt = np.asarray(a).shape[-1]
    assert fmt in (5, 8)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, fmt))
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1, fmt))
    assert a.shape == b.shape
    ov = np.empty((a.shape[0],), np.float64)
    grad = np.empty((a.shape[0], 2 * fmt), np.float64)
    fn = lib().oracle_riou_aligned_grad_fd if fmt == 5 else lib().oracle_riou_aligned_grad_fd8
    fn(_p(a), _p(b), C.c_int(a.shape[0]), C.c_int({"iou": 0, "iof": 1, "iof_b": 2}[mode]),
                                      C.c_double(step), _p(ov), _p(grad))
    return ov, grad


def hbb_overlaps(a, b, mode="iou", plus_one=True):
    a, b = _f32(a, 4), _f32(b, 4)
    out = np.empty((a.shape[0], b.shape[0]), np.float64)
    lib().oracle_hbb_overlaps(_p(a), C.c_int(a.shape[0]), _p(b), C.c_int(b.shape[0]),
                              C.c_int(MODE_IOF if mode == "iof" else MODE_IOU), C.c_int(int(plus_one)), _p(out))
    return out


def _nms_args(boxes, scores, groups, thr):
    boxes = np.asarray(boxes, dtype=np.float32)
    fmt = boxes.shape[-1]
    boxes = _f32(boxes, fmt)
    n = boxes.shape[0]
    scores = _f32(scores).reshape(-1)
    assert scores.shape[0] == n
    g = None if groups is None else np.ascontiguousarray(np.asarray(groups, dtype=np.int32).reshape(-1))
    thr = np.ascontiguousarray(np.atleast_1d(np.asarray(thr, dtype=np.float64)))
    return boxes, fmt, n, scores, g, thr


def nms(boxes, scores, thr, groups=None, cmp_ge=False, plus_one=True, margin=1e-6):
    """Greedy (batched) NMS.  Returns (keep ascending original idx, #pairs within margin of thr)."""
    boxes, fmt, n, scores, g, thr = _nms_args(boxes, scores, groups, thr)
    keep = np.empty((max(n, 1),), np.int64)
    near = C.c_int64(0)
    k = lib().oracle_nms(_p(boxes), C.c_int(fmt), _p(scores), _p(g) if g is not None else None, C.c_int(n),
                         _p(thr), C.c_int(thr.shape[0]), C.c_int(int(cmp_ge)), C.c_int(int(plus_one)),
                         C.c_double(margin), _p(keep), C.byref(near))
    return keep[:k].copy(), int(near.value)


def nms_verify(boxes, scores, thr, keep, groups=None, cmp_ge=False, plus_one=True, margin=1e-6):
    """Check `keep` against the greedy definition with a +-margin band.  -> (violations, near pairs)."""
    boxes, fmt, n, scores, g, thr = _nms_args(boxes, scores, groups, thr)
    keep = np.ascontiguousarray(np.asarray(keep, dtype=np.int64).reshape(-1))
    near = C.c_int64(0)
    bad = lib().oracle_nms_verify(_p(boxes), C.c_int(fmt), _p(scores), _p(g) if g is not None else None,
                                  C.c_int(n), _p(thr), C.c_int(thr.shape[0]), C.c_int(int(cmp_ge)),
                                  C.c_int(int(plus_one)), C.c_double(margin), _p(keep),
                                  C.c_int(keep.shape[0]), C.byref(near))
    return int(bad), int(near.value)


def roi_align_fwd(feat_nhwc, rois, scale, out_size, sample_num, variant, want_touched=False, unit_weights=False):
    """feat (N,H,W,C) f32, rois (K,5|6) -> (K,ph,pw,C) f64 [, touched (N,H,W) bool].
    unit_weights=True (error-bound mode for the tests): every accepted tap weighs 1 -> sum_taps x_t / count."""
    lib().oracle_roi_set_unit_weights(C.c_int(int(unit_weights)))
    feat = np.ascontiguousarray(np.asarray(feat_nhwc, dtype=np.float32))
    N, H, W, Cc = feat.shape
    rois = np.asarray(rois, dtype=np.float32)
    fmt = rois.shape[-1]
    rois = _f32(rois, fmt)
    ph, pw = out_size
    out = np.empty((rois.shape[0], ph, pw, Cc), np.float64)
    touched = np.zeros((N, H, W), np.uint8) if want_touched else None
    lib().oracle_roi_align_fwd(_p(feat), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(Cc), _p(rois),
                               C.c_int(fmt), C.c_int(rois.shape[0]), C.c_double(scale), C.c_int(ph),
                               C.c_int(pw), C.c_int(sample_num), C.c_int(variant), _p(out),
                               _p(touched) if want_touched else None)
    lib().oracle_roi_set_unit_weights(C.c_int(0))
    return (out, touched.astype(bool)) if want_touched else out


def roi_align_bwd(grad_out, feat_shape_nhwc, rois, scale, sample_num, variant, unit_weights=False):
    """grad_out (K,ph,pw,C) f32 -> grad_feat (N,H,W,C) f64.  unit_weights: see roi_align_fwd."""
    lib().oracle_roi_set_unit_weights(C.c_int(int(unit_weights)))
    go = np.ascontiguousarray(np.asarray(grad_out, dtype=np.float32))
    K, ph, pw, Cc = go.shape
    N, H, W, C2 = feat_shape_nhwc
    assert C2 == Cc
    rois = np.asarray(rois, dtype=np.float32)
    fmt = rois.shape[-1]
    rois = _f32(rois, fmt)
    gf = np.zeros((N, H, W, Cc), np.float64)
    lib().oracle_roi_align_bwd(_p(go), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(Cc), _p(rois),
                               C.c_int(fmt), C.c_int(K), C.c_double(scale), C.c_int(ph), C.c_int(pw),
                               C.c_int(sample_num), C.c_int(variant), _p(gf))
    lib().oracle_roi_set_unit_weights(C.c_int(0))
    return gf


def map_roi_levels(rois5, finest_scale=56, num_levels=4):
    r = _f32(rois5, 5)
    out = np.empty((r.shape[0],), np.int32)
    lib().oracle_map_roi_levels(_p(r), C.c_int(r.shape[0]), C.c_double(finest_scale), C.c_int(num_levels), _p(out))
    return out


def max_iou_assign_wrt_overlaps(overlaps, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, gt_max_assign_all=True,
                                gt_labels=None):
    """numpy restatement of MaxIoUAssigner.assign_wrt_overlaps
    (mmdet/core/bbox/assigners/max_iou_assigner.py:122-195), step by step; pinned to outputs of the reference's own
    class in tests/golden/golden_assign_v1.npz.  overlaps (k, n), -1 entries = ignored.
    -> (gt_inds (n,) int64, max_overlaps (n,), labels (n,) int64 | None)"""
    ov = np.asarray(overlaps)
    k, n = ov.shape
    gt_inds = np.full((n,), -1, np.int64)                                   # :137-139
    if k == 0 or n == 0:                                                    # :141-153
        if k == 0:
            gt_inds[:] = 0
        return gt_inds, np.zeros((n,), ov.dtype), (None if gt_labels is None else np.zeros((n,), np.int64))
    max_ov, argmax_ov = ov.max(0), ov.argmax(0)                             # :155  (first index on ties)
    gt_max, gt_argmax = ov.max(1), ov.argmax(1)                             # :158
    if isinstance(neg_iou_thr, float):                                      # :161-163
        gt_inds[(max_ov >= 0) & (max_ov < neg_iou_thr)] = 0
    elif isinstance(neg_iou_thr, tuple):                                    # :164-167
        gt_inds[(max_ov >= neg_iou_thr[0]) & (max_ov < neg_iou_thr[1])] = 0
    pos = max_ov >= pos_iou_thr                                             # :170-171
    gt_inds[pos] = argmax_ov[pos] + 1
    for i in range(k):                                                      # :174-182
        if gt_max[i] >= min_pos_iou:
            if gt_max_assign_all:
                gt_inds[ov[i, :] == gt_max[i]] = i + 1
            else:
                gt_inds[gt_argmax[i]] = i + 1
    labels = None
    if gt_labels is not None:                                               # :184-190
        labels = np.zeros((n,), np.int64)
        p = gt_inds > 0
        labels[p] = np.asarray(gt_labels)[gt_inds[p] - 1]
    return gt_inds, max_ov, labels


def max_iou_assign(bboxes, gt_bboxes, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, gt_max_assign_all=True,
                   ignore_iof_thr=-1, ignore_wrt_candidates=True, gt_bboxes_ignore=None, gt_labels=None):
    """MaxIoUAssigner.assign (max_iou_assigner.py:99-120) on the float64 oracle overlaps; boxes (n,4|5|8)."""
    b, g = np.asarray(bboxes, np.float32), np.asarray(gt_bboxes, np.float32)
    fmt = g.shape[-1]
    over = (lambda x, y, mode="iou": hbb_overlaps(x, y, mode)) if fmt == 4 else (lambda x, y, mode="iou": riou_matrix(x, y, mode))
    ov = over(g, b)                                                         # :102
    if ignore_iof_thr > 0 and gt_bboxes_ignore is not None and len(gt_bboxes_ignore) and len(b):   # :104-113
        ig = np.asarray(gt_bboxes_ignore, np.float32)
        if ignore_wrt_candidates:
            ign_max = over(b, ig, "iof").max(1)
        else:
            ign_max = over(ig, b, "iof").max(0)
        ov[:, ign_max > ignore_iof_thr] = -1
    return max_iou_assign_wrt_overlaps(ov, pos_iou_thr, neg_iou_thr, min_pos_iou, gt_max_assign_all, gt_labels)
