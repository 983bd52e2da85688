"""Thin tensor-level wrappers over the C ABI (one function per entry point of include/aidet_b200.h).

Inputs are CUDA float32 tensors; nothing here touches the CPU or the oracle.
"""
import ctypes as C

import torch

from .. import _lib as L


def _f32c(t, cols, name):
    L.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.dim() != 2 or t.size(1) != cols:
        raise ValueError("%s must have shape (n, %d), got %s" % (name, cols, tuple(t.shape)))
    return t


def riou_matrix(a, b, mode="iou", out=None):
    """(m,fmt) x (n,fmt) -> (m,n) float32 overlap matrix, fmt 5 (cx,cy,w,h,theta) or 8 (x1..y4)."""
    assert mode in ("iou", "iof")
    fmt = a.size(-1)
    a, b = _f32c(a, fmt, "a"), _f32c(b, fmt, "b")
    if a.device != b.device:
        raise ValueError("a and b must be on the same device")
    m, n = a.size(0), b.size(0)
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    if m == 0 or n == 0:
        return out
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (m, n)
    dev = a.device.index
    lib = L.lib()
    ws_bytes = lib.aidet_riou_workspace_bytes(m, n, fmt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(dev):
        L.check(lib.aidet_riou_matrix_f32(L.dptr(a), m, L.dptr(b), n, fmt, L.MODE_IOF if mode == "iof" else L.MODE_IOU,
                                          L.dptr(out), out.stride(0), L.dptr(ws), ws_bytes, dev, L.stream_ptr(dev)),
                "aidet_riou_matrix_f32")
    return out


def riou_matrix_multi(a, b, dst_ptrs, ld, mode="iou"):
    """Overlap matrix of a (m,fmt) x b (n,fmt) stored to every raw device pointer in `dst_ptrs` (row stride `ld`
    elements): the local block and the peer-mapped blocks of the other GPUs (see aidet_b200.sharded)."""
    assert mode in ("iou", "iof")
    fmt = a.size(-1)
    a, b = _f32c(a, fmt, "a"), _f32c(b, fmt, "b")
    m, n = a.size(0), b.size(0)
    if m == 0 or n == 0:
        return
    dev = a.device.index
    lib = L.lib()
    ws_bytes = lib.aidet_riou_workspace_bytes(m, n, fmt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    ptrs = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
    with torch.cuda.device(dev):
        L.check(lib.aidet_riou_matrix_multi_f32(L.dptr(a), m, L.dptr(b), n, fmt,
                                                L.MODE_IOF if mode == "iof" else L.MODE_IOU, ptrs, len(dst_ptrs),
                                                int(ld), L.dptr(ws), ws_bytes, dev, L.stream_ptr(dev)),
                "aidet_riou_matrix_multi_f32")


def riou_matrix_mcast(a, b, mc_ptr, ld, mode="iou"):
    """Overlap matrix of a (m,fmt) x b (n,fmt) stored ONCE to the NVSwitch multicast address `mc_ptr` (row stride
    `ld` elements); the switch replicates every store into all GPUs of the multicast group (see aidet_b200.sharded)."""
    assert mode in ("iou", "iof")
    fmt = a.size(-1)
    a, b = _f32c(a, fmt, "a"), _f32c(b, fmt, "b")
    m, n = a.size(0), b.size(0)
    if m == 0 or n == 0:
        return
    dev = a.device.index
    lib = L.lib()
    ws_bytes = lib.aidet_riou_workspace_bytes(m, n, fmt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(dev):
        L.check(lib.aidet_riou_matrix_mcast_f32(L.dptr(a), m, L.dptr(b), n, fmt,
                                                L.MODE_IOF if mode == "iof" else L.MODE_IOU, C.c_void_p(int(mc_ptr)),
                                                int(ld), L.dptr(ws), ws_bytes, dev, L.stream_ptr(dev)),
                "aidet_riou_matrix_mcast_f32")


def riou_aligned(a, b, mode="iou"):
    assert mode in ("iou", "iof")
    fmt = a.size(-1)
    a, b = _f32c(a, fmt, "a"), _f32c(b, fmt, "b")
    if a.shape != b.shape:
        raise ValueError("aligned overlaps need equal shapes, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    n = a.size(0)
    out = torch.empty((n,), dtype=torch.float32, device=a.device)
    if n == 0:
        return out
    dev = a.device.index
    with torch.cuda.device(dev):
        L.check(L.lib().aidet_riou_aligned_f32(L.dptr(a), L.dptr(b), n, fmt,
                                               L.MODE_IOF if mode == "iof" else L.MODE_IOU, L.dptr(out), dev,
                                               L.stream_ptr(dev)), "aidet_riou_aligned_f32")
    return out


def riou_aligned_grad(a, b, grad_ov=None, mode="iou", want_b=True):
    """Backward of `riou_aligned` for theta-OBB (n,5) or convex point-OBB (n,8) pairs: (ov (n,), grad_a (n,fmt),
    grad_b (n,fmt) | None), the gradients scaled by `grad_ov` (n,) when given.  aidet_riou_aligned_grad_f32."""
    assert mode in ("iou", "iof")
    fmt = a.size(-1)
    if fmt not in (5, 8):
        raise ValueError("riou_aligned_grad takes (n, 5) or (n, 8) boxes, got %s" % (tuple(a.shape),))
    a, b = _f32c(a, fmt, "a"), _f32c(b, fmt, "b")
    if a.shape != b.shape:
        raise ValueError("aligned overlaps need equal shapes, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    n = a.size(0)
    ov = torch.empty((n,), dtype=torch.float32, device=a.device)
    ga = torch.empty((n, fmt), dtype=torch.float32, device=a.device)
    gb = torch.empty((n, fmt), dtype=torch.float32, device=a.device) if want_b else None
    if n == 0:
        return ov, ga, gb
    if grad_ov is not None:
        L.require_cuda(grad_ov, "grad_ov")
        grad_ov = grad_ov.float().contiguous().view(-1)
        assert grad_ov.numel() == n
    dev = a.device.index
    with torch.cuda.device(dev):
        L.check(L.lib().aidet_riou_aligned_grad_f32(L.dptr(a), L.dptr(b), n, fmt,
                                                    L.MODE_IOF if mode == "iof" else L.MODE_IOU, L.dptr(grad_ov),
                                                    L.dptr(ov), L.dptr(ga), L.dptr(gb), dev, L.stream_ptr(dev)),
                "aidet_riou_aligned_grad_f32")
    return ov, ga, gb


def _neg_bounds(neg_iou_thr):
    """(lo, hi) of max_iou_assigner.py:161-167: a float thr means [0, thr), a pair [thr[0], thr[1]), else none."""
    if isinstance(neg_iou_thr, float):
        return 0.0, neg_iou_thr
    if isinstance(neg_iou_thr, tuple):
        assert len(neg_iou_thr) == 2
        return float(neg_iou_thr[0]), float(neg_iou_thr[1])
    return float("inf"), float("inf")


def max_iou_assign(gts, bboxes, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, gt_max_assign_all=True, gt_ignore=None,
                   ignore_iof_thr=-1.0, ignore_wrt_candidates=True, gt_labels=None):
    """Fused overlap + MaxIoUAssigner steps (aidet_max_iou_assign_f32).  gts (k,fmt), bboxes (n,fmt), k, n >= 1,
    fmt 4 | 5 | 8.  -> (gt_inds (n,) int64, max_overlaps (n,) float32, labels (n,) int64 | None)."""
    fmt = bboxes.size(-1)
    gts, bboxes = _f32c(gts, fmt, "gt_bboxes"), _f32c(bboxes, fmt, "bboxes")
    k, n = gts.size(0), bboxes.size(0)
    assert k > 0 and n > 0
    k_ign = 0
    if gt_ignore is not None and gt_ignore.numel() > 0 and ignore_iof_thr > 0:
        gt_ignore = _f32c(gt_ignore, fmt, "gt_bboxes_ignore")
        k_ign = gt_ignore.size(0)
    else:
        gt_ignore = None
    if gt_labels is not None:
        L.require_cuda(gt_labels, "gt_labels")
        gt_labels = gt_labels.long().contiguous()
        assert gt_labels.numel() == k
    dev = bboxes.device
    gt_inds = torch.empty((n,), dtype=torch.long, device=dev)
    max_ov = torch.empty((n,), dtype=torch.float32, device=dev)
    labels = torch.empty((n,), dtype=torch.long, device=dev) if gt_labels is not None else None
    lo, hi = _neg_bounds(neg_iou_thr)
    lib = L.lib()
    ws_bytes = lib.aidet_assign_workspace_bytes(k, n, k_ign, fmt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev.index):
        L.check(lib.aidet_max_iou_assign_f32(L.dptr(gts), k, L.dptr(bboxes), n, fmt, L.dptr(gt_ignore), k_ign,
                                             float(ignore_iof_thr), int(bool(ignore_wrt_candidates)),
                                             float(pos_iou_thr), lo, hi, float(min_pos_iou),
                                             int(bool(gt_max_assign_all)), L.dptr(gt_labels), L.dptr(gt_inds),
                                             L.dptr(max_ov), L.dptr(labels), L.dptr(ws), ws_bytes, dev.index,
                                             L.stream_ptr(dev.index)), "aidet_max_iou_assign_f32")
    return gt_inds, max_ov, labels


def assign_wrt_overlaps(overlaps, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, gt_max_assign_all=True, gt_labels=None):
    """MaxIoUAssigner steps on a given (k, n) overlap matrix, k, n >= 1 (aidet_assign_wrt_overlaps_f32)."""
    L.require_cuda(overlaps, "overlaps")
    assert overlaps.dim() == 2
    ov = overlaps.float()
    if ov.stride(1) != 1:
        ov = ov.contiguous()
    k, n = ov.shape
    assert k > 0 and n > 0
    if gt_labels is not None:
        L.require_cuda(gt_labels, "gt_labels")
        gt_labels = gt_labels.long().contiguous()
        assert gt_labels.numel() == k
    dev = ov.device
    gt_inds = torch.empty((n,), dtype=torch.long, device=dev)
    max_ov = torch.empty((n,), dtype=torch.float32, device=dev)
    labels = torch.empty((n,), dtype=torch.long, device=dev) if gt_labels is not None else None
    lo, hi = _neg_bounds(neg_iou_thr)
    lib = L.lib()
    ws_bytes = lib.aidet_assign_workspace_bytes(k, n, 0, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev.index):
        L.check(lib.aidet_assign_wrt_overlaps_f32(L.dptr(ov), k, n, ov.stride(0), float(pos_iou_thr), lo, hi,
                                                  float(min_pos_iou), int(bool(gt_max_assign_all)),
                                                  L.dptr(gt_labels), L.dptr(gt_inds), L.dptr(max_ov), L.dptr(labels),
                                                  L.dptr(ws), ws_bytes, dev.index, L.stream_ptr(dev.index)),
                "aidet_assign_wrt_overlaps_f32")
    return gt_inds, max_ov, labels


_THR_CACHE = {}


def _scalar_thr(value, device):
    """(1,) float32 device tensor holding a threshold; cached so a call does not launch a fill kernel."""
    key = (device.type, device.index, value)
    t = _THR_CACHE.get(key)
    if t is None:
        if len(_THR_CACHE) > 256:
            _THR_CACHE.clear()
        t = torch.full((1,), value, dtype=torch.float32, device=device)
        _THR_CACHE[key] = t
    return t


_NMS_WS = {}
_COUNT_SLOTS = {}


def _count_slot(device):
    """(pinned int32 tensor (1,), ctypes view of it) per device and stream: where a synchronous call receives its count."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    slot = _COUNT_SLOTS.get(key)
    if slot is None:
        if len(_COUNT_SLOTS) > 64:
            _COUNT_SLOTS.clear()
        t = torch.empty((1,), dtype=torch.int32, pin_memory=True)
        slot = (t, C.c_int.from_address(t.data_ptr()))
        _COUNT_SLOTS[key] = slot
    return slot


def _nms_workspace(device, nbytes):
    """Per-device scratch that grows on demand and is reused by every call on that device's current stream (the
    reference cudaMallocs and frees its mask on every call, nms_kernel.cu:94,133).  Calls on the same stream are
    ordered, so sharing is safe there; a caller that runs NMS on several streams at once passes its own `workspace`."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _NMS_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if len(_NMS_WS) > 64:
            _NMS_WS.clear()
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _NMS_WS[key] = ws
    return ws


def nms_batched(boxes, scores, group_ids=None, iou_thr=0.5, n_groups=None, cmp_ge=False, plus_one=False, sync=True,
                workspace=None, keep_fill=None):
    """Batched greedy NMS.  boxes (n,4|5|8), scores (n,), group_ids (n,) int or None.

    iou_thr: float, or a (n_groups,) tensor / sequence of per-group thresholds.
    Returns keep (k,) int64, ascending original index.  sync=False: no host read-back -- returns (keep (n,) whose first
    n_keep entries are valid, n_keep (1,) int32 device tensor); the call is then capturable in a CUDA graph (pass
    n_groups and a float / tensor threshold so that nothing else touches the host).
    workspace: optional uint8 CUDA tensor of >= aidet_nms_workspace_bytes + 128 bytes owned by the caller; by default a
    cached per-(device, stream) buffer is reused, so a call allocates nothing but its two result tensors.
    keep_fill: with sync=False, the value the entries of `keep` past n_keep hold (the kernels only write the first n_keep);
    e.g. n, so that `mask = zeros(n + 1); mask[keep] = True` turns the result into a keep mask without reading n_keep.
    """
    fmt = boxes.size(-1)
    boxes = _f32c(boxes, fmt, "boxes")
    L.require_cuda(scores, "scores")
    n = boxes.size(0)
    scores = scores.reshape(-1).float().contiguous()
    if scores.numel() != n:
        raise ValueError("scores must have one entry per box")
    device = boxes.device
    if n == 0:
        empty = torch.zeros((0,), dtype=torch.long, device=device)
        return empty if sync else (empty, torch.zeros((1,), dtype=torch.int32, device=device))
    if group_ids is None:
        n_groups = 1
        gids = None
    else:
        gids = group_ids.reshape(-1).to(device=device, dtype=torch.int32).contiguous()
        if gids.numel() != n:
            raise ValueError("group_ids must have one entry per box")
        if n_groups is None:
            n_groups = int(gids.max().item()) + 1
    if isinstance(iou_thr, torch.Tensor):
        thr = iou_thr.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    elif isinstance(iou_thr, (list, tuple)):
        thr = torch.tensor(list(iou_thr), dtype=torch.float32, device=device)
    else:
        thr = _scalar_thr(float(iou_thr), device)
    if thr.numel() not in (1, n_groups):
        raise ValueError("iou_thr must be a scalar or have n_groups=%d entries" % n_groups)
    dev = device.index
    lib = L.lib()
    keep = (torch.empty((n,), dtype=torch.long, device=device) if keep_fill is None
            else torch.full((n,), int(keep_fill), dtype=torch.long, device=device))
    # sync=True: the kernels write the count straight into pinned host memory (device-accessible under unified addressing)
    # and the host polls it -- a few microseconds instead of the ~15 of a blocking device-to-host copy of 4 bytes
    slot = _count_slot(device) if sync else None
    if slot is not None:
        slot[1].value = -1
    n_keep = slot[0] if slot is not None else torch.empty((1,), dtype=torch.int32, device=device)
    with torch.cuda.device(dev):
        ws_bytes = lib.aidet_nms_workspace_bytes(n, n_groups, fmt)
        if ws_bytes == 0:
            L.check(-2, "aidet_nms_workspace_bytes")
        ws = workspace if workspace is not None else _nms_workspace(device, ws_bytes + 128)
        if ws.numel() < ws_bytes + 128 or ws.device != device or ws.dtype != torch.uint8:
            raise ValueError("workspace must be a uint8 tensor of >= %d bytes on %s" % (ws_bytes + 128, device))
        base = ws.data_ptr()
        aligned = (base + 127) // 128 * 128
        L.check(lib.aidet_nms_batched_f32(L.dptr(boxes), fmt, L.dptr(scores), L.dptr(gids), n, L.dptr(thr),
                                          thr.numel(), n_groups, L.CMP_GE if cmp_ge else L.CMP_GT,
                                          int(bool(plus_one)), L.dptr(keep), L.dptr(n_keep), C.c_void_p(aligned),
                                          ws_bytes, dev, L.stream_ptr(dev)), "aidet_nms_batched_f32")
    if not sync:
        return keep, n_keep
    cell = slot[1]
    k = cell.value
    spins = 0
    while k < 0 and spins < 4096:                            # ~0.2 ms of polling covers every per-image config
        k = cell.value
        spins += 1
    if k < 0:                                                # long call (or a failed kernel: the synchronize raises)
        torch.cuda.current_stream(device).synchronize()
        k = cell.value
        if k < 0:
            raise RuntimeError("aidet_nms_batched_f32: the kernels left no keep count")
    return keep[:k]


def soft_nms_batched(dets, group_ids=None, iou_thr=0.3, method=1, sigma=0.5, min_score=1e-3, n_groups=None):
    """Soft-NMS of dets (n,5) [x1,y1,x2,y2,score] per group (one CTA per group, all groups in one launch).

    Returns (rows (k,6) [x1,y1,x2,y2,decayed score,original index], counts (n_groups,) int64): the rows of
    group g are the `counts[g]` rows after those of the groups before it, in the reference's selection order
    (nms_cpu.cpp:190-199)."""
    dets = _f32c(dets, 5, "dets")
    n = dets.size(0)
    device = dets.device
    if group_ids is None:
        n_groups = 1
        order = torch.arange(n, device=device)
        offsets = torch.tensor([0, n], dtype=torch.int32, device=device)
        max_group = n
    else:
        gids = group_ids.reshape(-1).to(device=device, dtype=torch.long)
        if n_groups is None:
            n_groups = int(gids.max().item()) + 1 if n else 1
        order = torch.argsort(gids, stable=True)                     # groups contiguous, original order inside
        cnt = torch.bincount(gids, minlength=n_groups)
        offsets = torch.zeros(n_groups + 1, dtype=torch.int32, device=device)
        offsets[1:] = torch.cumsum(cnt, 0).int()
        max_group = int(cnt.max().item()) if n else 0
    rows = torch.cat([dets[order], order.float()[:, None]], dim=1).contiguous()
    n_out = torch.zeros(n_groups, dtype=torch.int32, device=device)
    if n:
        dev = device.index
        with torch.cuda.device(dev):
            L.check(L.lib().aidet_soft_nms_f32(L.dptr(rows), L.dptr(offsets), n_groups, max_group, float(iou_thr),
                                               int(method), float(sigma), float(min_score), L.dptr(n_out), dev,
                                               L.stream_ptr(dev)), "aidet_soft_nms_f32")
    counts = n_out.long()
    pos = torch.arange(n, device=device)
    grp = torch.bucketize(pos, offsets[1:].long(), right=True).clamp(max=n_groups - 1)
    keep = (pos - offsets.long()[grp]) < counts[grp]
    return rows[keep], counts


def _level_tables(tensors):
    n = len(tensors)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    Hs = (C.c_int * n)(*[t.size(1) for t in tensors])
    Ws = (C.c_int * n)(*[t.size(2) for t in tensors])
    return ptrs, Hs, Ws


def rroi_align_forward(feats_nhwc, rois, scales, out_size, sample_num, variant, roi_level=None, out=None):
    """feats_nhwc: list of (N,H_l,W_l,C) contiguous float32; rois (K,5|6) -> (K,ph,pw,C)."""
    ph, pw = out_size
    f0 = feats_nhwc[0]
    N, C_ = f0.size(0), f0.size(3)
    for t in feats_nhwc:
        L.require_cuda(t, "features")
        assert t.dtype == torch.float32 and t.is_contiguous() and t.size(0) == N and t.size(3) == C_
    rois = _f32c(rois, rois.size(-1), "rois")
    K = rois.size(0)
    if out is None:
        out = torch.empty((K, ph, pw, C_), dtype=torch.float32, device=f0.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape == (K, ph, pw, C_)
    if K == 0:
        return out
    lvl = None if roi_level is None else roi_level.to(device=f0.device, dtype=torch.int32).contiguous()
    ptrs, Hs, Ws = _level_tables(feats_nhwc)
    sc = (C.c_float * len(scales))(*[float(s) for s in scales])
    dev = f0.device.index
    with torch.cuda.device(dev):
        L.check(L.lib().aidet_rroi_align_fwd_f32(ptrs, Hs, Ws, sc, len(feats_nhwc), N, C_, L.dptr(rois), rois.size(1),
                                                 L.dptr(lvl), K, ph, pw, int(sample_num), int(variant), L.dptr(out),
                                                 dev, L.stream_ptr(dev)), "aidet_rroi_align_fwd_f32")
    return out


def rroi_align_backward(grad_out_nhwc, grad_feats_nhwc, rois, scales, sample_num, variant, roi_level=None):
    """Accumulates into the (pre-zeroed) grad_feats_nhwc list; grad_out (K,ph,pw,C) contiguous."""
    g0 = grad_feats_nhwc[0]
    N, C_ = g0.size(0), g0.size(3)
    L.require_cuda(grad_out_nhwc, "grad_output")
    assert grad_out_nhwc.dtype == torch.float32 and grad_out_nhwc.is_contiguous()
    K, ph, pw, C2 = grad_out_nhwc.shape
    assert C2 == C_
    for t in grad_feats_nhwc:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.size(0) == N and t.size(3) == C_
    if K == 0:
        return
    rois = _f32c(rois, rois.size(-1), "rois")
    lvl = None if roi_level is None else roi_level.to(device=g0.device, dtype=torch.int32).contiguous()
    ptrs, Hs, Ws = _level_tables(grad_feats_nhwc)
    sc = (C.c_float * len(scales))(*[float(s) for s in scales])
    dev = g0.device.index
    with torch.cuda.device(dev):
        L.check(L.lib().aidet_rroi_align_bwd_f32(L.dptr(grad_out_nhwc), ptrs, Hs, Ws, sc, len(grad_feats_nhwc), N, C_,
                                                 L.dptr(rois), rois.size(1), L.dptr(lvl), K, ph, pw, int(sample_num),
                                                 int(variant), dev, L.stream_ptr(dev)), "aidet_rroi_align_bwd_f32")


def rroi_align_backward_gather(grad_out_nhwc, grad_feats_nhwc, rois, scales, sample_num, variant, roi_level=None,
                               workspace=None, deterministic=False):
    """Gather backward: OVERWRITES the grad_feats_nhwc list (no zero-fill needed, no atomics on the maps).
    deterministic=True buckets the taps with a stable radix sort (bit-reproducible gradients).

    Needs sample_num > 0 and C % 4 == 0; `workspace` (uint8 CUDA tensor) is allocated when not given.
    """
    g0 = grad_feats_nhwc[0]
    N, C_ = g0.size(0), g0.size(3)
    L.require_cuda(grad_out_nhwc, "grad_output")
    assert grad_out_nhwc.dtype == torch.float32 and grad_out_nhwc.is_contiguous()
    K, ph, pw, C2 = grad_out_nhwc.shape
    assert C2 == C_ and sample_num > 0 and C_ % 4 == 0
    for t in grad_feats_nhwc:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.size(0) == N and t.size(3) == C_
    rois = _f32c(rois, rois.size(-1), "rois")
    lvl = None if roi_level is None else roi_level.to(device=g0.device, dtype=torch.int32).contiguous()
    ptrs, Hs, Ws = _level_tables(grad_feats_nhwc)
    sc = (C.c_float * len(scales))(*[float(s) for s in scales])
    dev = g0.device.index
    lib = L.lib()
    with torch.cuda.device(dev):
        ws_ptr, ws_bytes = C.c_void_p(0), 0
        if K > 0:
            ws_bytes = lib.aidet_rroi_align_bwd_workspace_bytes(Hs, Ws, len(grad_feats_nhwc), N, K, ph, pw, int(sample_num))
            if ws_bytes == 0:
                raise RuntimeError("aidet_rroi_align_bwd_workspace_bytes: problem too large for the gather path")
            if workspace is None or workspace.numel() < ws_bytes + 256:
                workspace = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=g0.device)
            ws_ptr = C.c_void_p((workspace.data_ptr() + 255) // 256 * 256)
        L.check(lib.aidet_rroi_align_bwd_gather_f32(L.dptr(grad_out_nhwc), ptrs, Hs, Ws, sc, len(grad_feats_nhwc), N, C_,
                                                    L.dptr(rois), rois.size(1), L.dptr(lvl), K, ph, pw,
                                                    int(sample_num), int(variant), int(bool(deterministic)), ws_ptr,
                                                    ws_bytes, dev,
                                                    L.stream_ptr(dev)), "aidet_rroi_align_bwd_gather_f32")
    return workspace
