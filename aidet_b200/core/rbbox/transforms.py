"""OBB codecs of mmdet/core/rbbox/transforms.py as batched tensor ops (they run on whatever device the inputs
live on, so decode -> rescale -> NMS stays on the GPU between the head and the kernels).

These define what (cx, cy, w, h, theta) and the 8-point order MEAN for the kernels:
    thetaobb2pointobb  transforms.py:45-55   cv2.boxPoints(((cx, cy), (w, h), theta * 180 / pi)) -- per box, numpy
    pointobb2bbox      transforms.py:57-71   axis-aligned envelope
    thetaobb_rescale   transforms.py:280-293 scale everything but theta (in place, like the reference)
    pointobb_rescale   transforms.py:295-306
    thetaobb2delta     transforms.py:321-353 proposal (x1,y1,x2,y2) + gt theta-OBB -> deltas, base angle -pi/2
    delta2thetaobb     transforms.py:356-395
    pointobb2delta     transforms.py:412-456
    delta2pointobb     transforms.py:458-505
    hobb2pointobb      transforms.py:137-163 H-OBB (first edge + height) -> corners
    hobb2delta / delta2hobb / hobb_rescale  transforms.py:522-600,308-319
    rbbox2result       transforms.py:615-633
    pointobb2thetaobb  transforms.py:30-43   int truncation + cv2.minAreaRect (host, numpy)
    thetaobb2hobb, pointobb_extreme_sort, pointobb_best_point_sort   transforms.py:73-135
    thetaobb_flip / pointobb_flip / hobb_flip                        transforms.py:191-275
    thetaobb_mapping(_back) / pointobb_mapping(_back) / hobb_mapping(_back)   transforms.py:398-409,507-519,602-612
    rbbox2roi          (new) per-image (n, 5|6) theta-OBB detections -> (K, 6) [batch, cx, cy, w, h, theta] for RoIAlignRotated
The reference's list-in / list-out helpers take one box; the forms here also take (..., 5) / (..., 8) batches.
Outputs are pinned to the reference's own functions by tests/golden/golden_rbbox_v1.npz (made by
tests/golden/make_golden_rbbox.py from /root/reference).
"""
import math

import numpy as np
import torch


def thetaobb2pointobb(thetaobb):
    """(..., 5) [cx, cy, w, h, theta(rad)] -> (..., 8) corners in cv2.boxPoints order (transforms.py:45-55).

    boxPoints: p0 = c - (cos, sin) w/2 + (-sin, cos) h/2 ... written out (checked against cv2 in the tests):
        p0 = (cx - s h - c w, cy + c h - s w), p1 = (cx + s h - c w, cy - c h - s w), p2 = 2 ctr - p0, p3 = 2 ctr - p1
    with c = cos(theta) / 2, s = sin(theta) / 2.  Lists / tuples / ndarrays go through cv2.boxPoints itself (float32
    inside, exactly the reference's values -- the point-sort helpers below compare coordinates for equality); one box
    returns a python list like the reference, a batch an ndarray."""
    if not isinstance(thetaobb, torch.Tensor):
        import cv2
        a = np.asarray(thetaobb, dtype=np.float64)
        flat = a.reshape(-1, 5)
        pts = np.stack([cv2.boxPoints(((b[0], b[1]), (b[2], b[3]), b[4] * 180.0 / np.pi)).reshape(-1) for b in flat]) \
            if flat.shape[0] else np.zeros((0, 8), np.float32)
        return pts[0].tolist() if a.ndim == 1 else pts.reshape(a.shape[:-1] + (8,))
    cx, cy, w, h, th = thetaobb.unbind(-1)
    c, s = torch.cos(th) * 0.5, torch.sin(th) * 0.5
    x0 = cx - s * h - c * w
    y0 = cy + c * h - s * w
    x1 = cx + s * h - c * w
    y1 = cy - c * h - s * w
    return torch.stack([x0, y0, x1, y1, 2 * cx - x0, 2 * cy - y0, 2 * cx - x1, 2 * cy - y1], dim=-1)


def pointobb2bbox(pointobb):
    """(..., 8) -> (..., 4) [xmin, ymin, xmax, ymax] (transforms.py:57-71)."""
    single = not isinstance(pointobb, torch.Tensor)
    t = torch.as_tensor(np.asarray(pointobb, dtype=np.float64)) if single else pointobb
    xs, ys = t[..., 0::2], t[..., 1::2]
    out = torch.stack([xs.min(-1).values, ys.min(-1).values, xs.max(-1).values, ys.max(-1).values], dim=-1)
    return out.tolist() if single else out


def thetaobb_rescale(thetaobbs, scale_factor, reverse_flag=False):
    """Scale (..., 5k) theta-OBBs, theta untouched; IN PLACE and returned, as transforms.py:280-293."""
    keep = thetaobbs[..., 4::5].clone()
    if not reverse_flag:
        thetaobbs *= scale_factor
    else:
        thetaobbs /= scale_factor
    thetaobbs[..., 4::5] = keep
    return thetaobbs


def pointobb_rescale(pointobbs, scale_factor, reverse_flag=False):
    """transforms.py:295-306 (in place)."""
    if not reverse_flag:
        pointobbs *= scale_factor
    else:
        pointobbs /= scale_factor
    return pointobbs


def _proposal_cxcywh(p):
    return ((p[..., 0] + p[..., 2]) * 0.5, (p[..., 1] + p[..., 3]) * 0.5,
            p[..., 2] - p[..., 0] + 1.0, p[..., 3] - p[..., 1] + 1.0)


def thetaobb2delta(proposals, gt, means=(0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1)):
    """proposals (n,4) x1y1x2y2 (+1 sides), gt (n,5) -> (n,5) deltas; proposal angle is -pi/2 (transforms.py:321-353)."""
    assert proposals.size(0) == gt.size(0)
    proposals, gt = proposals.float(), gt.float()
    px, py, pw, ph = _proposal_cxcywh(proposals)
    gw, gh = gt[..., 2] + 1.0, gt[..., 3] + 1.0
    deltas = torch.stack([(gt[..., 0] - px) / pw, (gt[..., 1] - py) / ph, torch.log(gw / pw), torch.log(gh / ph),
                          gt[..., 4] + math.pi / 2.0], dim=-1)
    means = deltas.new_tensor(means).unsqueeze(0)
    stds = deltas.new_tensor(stds).unsqueeze(0)
    return deltas.sub_(means).div_(stds)


def delta2thetaobb(rois, deltas, means=(0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    """rois (n,4), deltas (n,5k) -> (n,5k) theta-OBBs (transforms.py:356-395; max_shape is ignored there too)."""
    k = deltas.size(1) // 5
    d = deltas * deltas.new_tensor(stds).repeat(1, k) + deltas.new_tensor(means).repeat(1, k)
    dx, dy, dw, dh, dth = d[:, 0::5], d[:, 1::5], d[:, 2::5], d[:, 3::5], d[:, 4::5]
    max_ratio = abs(math.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px, py, pw, ph = (v.unsqueeze(1) for v in _proposal_cxcywh(rois))
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + pw * dx, py + ph * dy
    gth = dth - math.pi / 2.0
    return torch.stack([gx, gy, gw, gh, gth], dim=-1).view_as(deltas)


def pointobb2delta(proposals, gt, means=(0,) * 8, stds=(1,) * 8):
    """proposals (n,4), gt (n,8) -> (n,8): corner offsets from the proposal's corners in units of its (+1) sides
    (transforms.py:412-456)."""
    assert proposals.size(0) == gt.size(0)
    proposals, gt = proposals.float(), gt.float()
    pw = proposals[..., 2] - proposals[..., 0] + 1.0
    ph = proposals[..., 3] - proposals[..., 1] + 1.0
    x1, y1, x2, y2 = proposals[..., 0], proposals[..., 1], proposals[..., 2], proposals[..., 3]
    deltas = torch.stack([(gt[..., 0] - x1) / pw, (gt[..., 1] - y1) / ph, (gt[..., 2] - x2) / pw, (gt[..., 3] - y1) / ph,
                          (gt[..., 4] - x2) / pw, (gt[..., 5] - y2) / ph, (gt[..., 6] - x1) / pw, (gt[..., 7] - y2) / ph],
                         dim=-1)
    means = deltas.new_tensor(means).unsqueeze(0)
    stds = deltas.new_tensor(stds).unsqueeze(0)
    return deltas.sub_(means).div_(stds)


def delta2pointobb(rois, deltas, means=(0,) * 8, stds=(1,) * 8, max_shape=None, wh_ratio_clip=16 / 1000):
    """rois (n,4), deltas (n,8k) -> (n,8k) point-OBBs (transforms.py:458-505)."""
    k = deltas.size(1) // 8
    d = deltas * deltas.new_tensor(stds).repeat(1, k) + deltas.new_tensor(means).repeat(1, k)
    pw = (rois[:, 2] - rois[:, 0] + 1.0).unsqueeze(1)
    ph = (rois[:, 3] - rois[:, 1] + 1.0).unsqueeze(1)
    x1, y1, x2, y2 = (rois[:, i].unsqueeze(1) for i in range(4))
    out = torch.stack([pw * d[:, 0::8] + x1, ph * d[:, 1::8] + y1, pw * d[:, 2::8] + x2, ph * d[:, 3::8] + y1,
                       pw * d[:, 4::8] + x2, ph * d[:, 5::8] + y2, pw * d[:, 6::8] + x1, ph * d[:, 7::8] + y2], dim=-1)
    return out.view_as(deltas)


def hobb2pointobb(hobb, truncate=None):
    """(..., 5) H-OBB [x1, y1, x2, y2, h] (first edge p1->p2 plus the box height) -> (..., 8) corners
    (transforms.py:137-163): p3 = p2 + h * (-cos a, sin a), p4 = p1 + h * (-cos a, sin a) with
    a = pi/2 - atan2(y2 - y1, x2 - x1).  The reference's list form truncates every coordinate to int (:161);
    that is kept for list / ndarray input (truncate=None -> True) and off for tensors (decoded predictions stay
    float until the NMS), override with `truncate`."""
    single = not isinstance(hobb, torch.Tensor)
    t = torch.as_tensor(np.asarray(hobb, dtype=np.float64)) if single else hobb
    x1, y1, x2, y2, h = t.unbind(-1)
    ang = math.pi / 2.0 - torch.atan2(y2 - y1, x2 - x1)
    dx, dy = h * torch.cos(ang), h * torch.sin(ang)
    out = torch.stack([x1, y1, x2, y2, x2 - dx, y2 + dy, x1 - dx, y1 + dy], dim=-1)
    if truncate if truncate is not None else single:
        out = out.trunc()
    return [int(v) for v in out.tolist()] if single and out.dim() == 1 else (out.tolist() if single else out)


def hobb_rescale(hobbs, scale_factor, reverse_flag=False):
    """transforms.py:308-319 (in place)."""
    if not reverse_flag:
        hobbs *= scale_factor
    else:
        hobbs /= scale_factor
    return hobbs


def hobb2delta(proposals, gt, means=(0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1)):
    """proposals (n,4), gt (n,5) [x1,y1,x2,y2,h] -> (n,5) deltas (transforms.py:522-556)."""
    assert proposals.size(0) == gt.size(0)
    proposals, gt = proposals.float(), gt.float()
    pw = proposals[..., 2] - proposals[..., 0] + 1.0
    ph = proposals[..., 3] - proposals[..., 1] + 1.0
    x1, y1, x2 = proposals[..., 0], proposals[..., 1], proposals[..., 2]
    deltas = torch.stack([(gt[..., 0] - x1) / pw, (gt[..., 1] - y1) / ph, (gt[..., 2] - x2) / pw, (gt[..., 3] - y1) / ph,
                          (gt[..., 4] + 1.0 - ph) / ph], dim=-1)
    means = deltas.new_tensor(means).unsqueeze(0)
    stds = deltas.new_tensor(stds).unsqueeze(0)
    return deltas.sub_(means).div_(stds)


def delta2hobb(rois, deltas, means=(0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    """rois (n,4), deltas (n,5k) -> (n,5k) H-OBBs (transforms.py:558-600)."""
    k = deltas.size(1) // 5
    d = deltas * deltas.new_tensor(stds).repeat(1, k) + deltas.new_tensor(means).repeat(1, k)
    max_ratio = abs(math.log(wh_ratio_clip))
    dh = d[:, 4::5].clamp(min=-max_ratio, max=max_ratio)
    pw = (rois[:, 2] - rois[:, 0] + 1.0).unsqueeze(1)
    ph = (rois[:, 3] - rois[:, 1] + 1.0).unsqueeze(1)
    x1, y1, x2 = (rois[:, i].unsqueeze(1) for i in range(3))
    out = torch.stack([pw * d[:, 0::5] + x1, ph * d[:, 1::5] + y1, pw * d[:, 2::5] + x2, ph * d[:, 3::5] + y1,
                       ph * dh + ph], dim=-1)
    return out.view_as(deltas)


def rbbox2result(rbboxes, labels, num_classes):
    """Detections -> list of per-class numpy arrays (transforms.py:615-633)."""
    if rbboxes.shape[0] == 0:
        return [np.zeros((0, 6), dtype=np.float32) for _ in range(num_classes - 1)]
    rbboxes = rbboxes.cpu().numpy()
    labels = labels.cpu().numpy()
    return [rbboxes[labels == i, :] for i in range(num_classes - 1)]


# ---------------------------------------------------------------- host-side converters (numpy / cv2, as the reference)
def _as_batch(x, d):
    """list / ndarray of one box or a batch -> ((n, d) float64 array, was it one box, leading shape)."""
    a = np.asarray(x, dtype=np.float64)
    assert a.shape[-1] == d, "expected (..., %d), got %s" % (d, a.shape)
    return a.reshape(-1, d), a.ndim == 1, a.shape[:-1]


def _ret(flat, one, lead):
    return flat[0].tolist() if one else flat.reshape(lead + flat.shape[1:])


def pointobb2thetaobb(pointobb):
    """(..., 8) corners -> (..., 5) [cx, cy, w, h, theta(rad)]: the minimum-area rectangle of the corner coordinates
    TRUNCATED to integers (transforms.py:30-43: np.int0 + cv2.minAreaRect; the angle convention is cv2's)."""
    import cv2
    pts, one, lead = _as_batch(pointobb, 8)
    out = np.zeros((pts.shape[0], 5), dtype=np.float64)
    for i, p in enumerate(pts.astype(np.intp).reshape(-1, 4, 2)):      # astype truncates towards zero like np.int0
        (x, y), (w, h), deg = cv2.minAreaRect(p.astype(np.int32))
        out[i] = (x, y, w, h, deg / 180.0 * np.pi)
    return _ret(out, one, lead)


def _roll_points(pts, k):
    """Cyclic left shift of every (8,) row of pts (n, 8) by k[i] POINTS: row i starts with its point k[i]."""
    idx = (np.arange(8)[None, :] + 2 * k[:, None]) % 8
    return np.take_along_axis(pts, idx, axis=1)


def pointobb_extreme_sort(pointobb):
    """Start every quad at its top point (smallest y; of two equally high top points the LEFT one), keeping the cyclic
    order (transforms.py:93-115)."""
    pts, one, lead = _as_batch(pointobb, 8)
    x, y = pts[:, 0::2], pts[:, 1::2]
    order = np.argsort(y, axis=1, kind="stable")
    first, second = order[:, 0], order[:, 1]
    rows = np.arange(pts.shape[0])
    tie = y[rows, first] == y[rows, second]
    top = np.where(tie & ~(x[rows, first] < x[rows, second]), second, first)
    return _ret(_roll_points(pts, top), one, lead)


def pointobb_best_point_sort(pointobb):
    """Of the four cyclic orders of a quad take the one closest (sum of squared corner distances) to its axis-aligned
    envelope walked as (xmin,ymin) (xmax,ymin) (xmax,ymax) (xmin,ymax); ties -> the smaller shift (transforms.py:118-135)."""
    pts, one, lead = _as_batch(pointobb, 8)
    x, y = pts[:, 0::2], pts[:, 1::2]
    xmin, xmax, ymin, ymax = x.min(1), x.max(1), y.min(1), y.max(1)
    ref = np.stack([xmin, ymin, xmax, ymin, xmax, ymax, xmin, ymax], 1)
    # np.roll(p, 2 r) of the reference shifts RIGHT by r points = starts at point (4 - r) % 4
    cand = np.stack([_roll_points(pts, np.full(pts.shape[0], (4 - r) % 4)) for r in range(4)], 1)      # (n, 4, 8)
    dist = ((cand - ref[:, None, :]) ** 2).sum(-1)
    best = np.argmin(dist, axis=1)                                   # first minimum, like argsort()[0] on 4 entries
    return _ret(cand[np.arange(pts.shape[0]), best], one, lead)


def thetaobb2hobb(thetaobb, pointobb_sort_fun=pointobb_best_point_sort):
    """(..., 5) theta-OBB -> (..., 5) H-OBB [x1, y1, x2, y2, h]: first and second corner of the sorted quad and the
    distance from the first to the fourth (transforms.py:73-90)."""
    tb, one, lead = _as_batch(thetaobb, 5)
    pts = np.asarray(thetaobb2pointobb(tb), dtype=np.float64).reshape(-1, 8)
    sp = np.asarray([pointobb_sort_fun(p.tolist()) for p in pts], dtype=np.float64).reshape(-1, 8) \
        if pointobb_sort_fun not in (pointobb_best_point_sort, pointobb_extreme_sort) else \
        np.asarray(pointobb_sort_fun(pts), dtype=np.float64).reshape(-1, 8)
    h = np.sqrt((sp[:, 6] - sp[:, 0]) ** 2 + (sp[:, 7] - sp[:, 1]) ** 2)
    return _ret(np.concatenate([sp[:, :4], h[:, None]], 1), one, lead)


# ---------------------------------------------------------------- flips and test-time-augmentation mappings
def _copy(a):
    return a.clone() if isinstance(a, torch.Tensor) else np.array(a, copy=True)


def thetaobb_flip(thetaobbs, img_shape):
    """Horizontal flip of (..., 5) theta-OBBs: x -> W - x - 1, w <-> h, theta -> -pi/2 - theta (transforms.py:191-203)."""
    assert thetaobbs.shape[-1] % 5 == 0
    flipped = _copy(thetaobbs)
    flipped[..., 0] = img_shape[1] - thetaobbs[..., 0] - 1
    flipped[..., 2] = thetaobbs[..., 3]
    flipped[..., 3] = thetaobbs[..., 2]
    flipped[..., 4] = -math.pi / 2.0 - thetaobbs[..., 4]
    return flipped


def pointobb_flip(pointobbs, img_shape):
    """Horizontal flip of (..., 8) point-OBBs: mirror x, swap corners 2 and 4 (orientation is kept), then restart every
    quad at its best point (transforms.py:205-240, the `pointobb_extreme_sort = False` branch)."""
    assert pointobbs.shape[-1] % 8 == 0
    tensor = isinstance(pointobbs, torch.Tensor)
    a = pointobbs.detach().cpu().numpy() if tensor else np.asarray(pointobbs)
    m = a.astype(np.float64, copy=True)
    m[..., 0::2] = img_shape[1] - m[..., 0::2] - 1
    m[..., [2, 3, 6, 7]] = m[..., [6, 7, 2, 3]]
    out = np.asarray(pointobb_best_point_sort(m.reshape(-1, 8)), dtype=np.float64).reshape(m.shape)
    return torch.as_tensor(out, dtype=pointobbs.dtype, device=pointobbs.device) if tensor else out


def hobb_flip(hobbs, img_shape):
    """Horizontal flip of (n, 5) H-OBBs by way of the corners: hobb2pointobb (integer corners) -> pointobb_flip ->
    pointobb2thetaobb -> thetaobb2hobb with the best-point order (transforms.py:243-275).  Always returns (n, 5)."""
    h = np.asarray(hobbs, dtype=np.float64)
    if h.ndim == 1:
        h = h[np.newaxis, ...]
    assert h.shape[-1] % 5 == 0
    pts = np.asarray([hobb2pointobb(b.tolist()) for b in h], dtype=np.float64).reshape(-1, 8)
    th = np.asarray(pointobb2thetaobb(pointobb_flip(pts, img_shape)), dtype=np.float64).reshape(-1, 5)
    return np.asarray(thetaobb2hobb(th, pointobb_best_point_sort), dtype=np.float64).reshape(-1, 5)


def _mapping(flip_fn, boxes, img_shape, scale_factor, flip):
    new = boxes * scale_factor               # every column, theta / h included, as the reference does
    return flip_fn(new, img_shape) if flip else new


def _mapping_back(flip_fn, boxes, img_shape, scale_factor, flip):
    new = flip_fn(boxes, img_shape) if flip else boxes
    return new / scale_factor


def thetaobb_mapping(thetaobbs, img_shape, scale_factor, flip):
    """Original image scale -> test scale (transforms.py:398-403).  NB the reference multiplies ALL five columns by
    `scale_factor`, the angle included; that is reproduced (goldens), not corrected."""
    return _mapping(thetaobb_flip, thetaobbs, img_shape, scale_factor, flip)


def thetaobb_mapping_back(thetaobbs, img_shape, scale_factor, flip):
    """Test scale -> original image scale (transforms.py:405-409)."""
    return _mapping_back(thetaobb_flip, thetaobbs, img_shape, scale_factor, flip)


def pointobb_mapping(pointobbs, img_shape, scale_factor, flip):
    """transforms.py:507-512."""
    return _mapping(pointobb_flip, pointobbs, img_shape, scale_factor, flip)


def pointobb_mapping_back(pointobbs, img_shape, scale_factor, flip):
    """transforms.py:514-518."""
    return _mapping_back(pointobb_flip, pointobbs, img_shape, scale_factor, flip)


def hobb_mapping(hobbs, img_shape, scale_factor, flip):
    """transforms.py:602-607."""
    return _mapping(hobb_flip, hobbs, img_shape, scale_factor, flip)


def hobb_mapping_back(hobbs, img_shape, scale_factor, flip):
    """transforms.py:609-613."""
    return _mapping_back(hobb_flip, hobbs, img_shape, scale_factor, flip)


def rbbox2roi(rbbox_list):
    """List of per-image (n_i, 5|6) theta-OBBs (a trailing score column is dropped) -> (K, 6) [batch index, cx, cy, w,
    h, theta], the RoI layout of `RoIAlignRotated` (the oriented twin of the reference's `bbox2roi`, which the OBB
    detectors call on horizontal proposals at mmdet/models/detectors/rbbox_cnn.py:177)."""
    rows = []
    for i, r in enumerate(rbbox_list):
        if r.size(0) > 0:
            rows.append(torch.cat([r.new_full((r.size(0), 1), i), r[:, :5]], dim=-1))
        else:
            rows.append(r.new_zeros((0, 6)))
    return torch.cat(rows, 0)
