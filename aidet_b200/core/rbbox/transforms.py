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
The reference's list-in / list-out helpers take one box; the tensor forms here take (..., 5) / (..., 8).
"""
import math

import numpy as np
import torch


def thetaobb2pointobb(thetaobb):
    """(..., 5) [cx, cy, w, h, theta(rad)] -> (..., 8) corners in cv2.boxPoints order (transforms.py:45-55).

    boxPoints: p0 = c - (cos, sin) w/2 + (-sin, cos) h/2 ... written out (checked against cv2 in the tests):
        p0 = (cx - s h - c w, cy + c h - s w), p1 = (cx + s h - c w, cy - c h - s w), p2 = 2 ctr - p0, p3 = 2 ctr - p1
    with c = cos(theta) / 2, s = sin(theta) / 2.  Lists / tuples / ndarrays of one box return a python list like
    the reference."""
    single = not isinstance(thetaobb, torch.Tensor)
    t = torch.as_tensor(np.asarray(thetaobb, dtype=np.float64)) if single else thetaobb
    cx, cy, w, h, th = t.unbind(-1)
    c, s = torch.cos(th) * 0.5, torch.sin(th) * 0.5
    x0 = cx - s * h - c * w
    y0 = cy + c * h - s * w
    x1 = cx + s * h - c * w
    y1 = cy - c * h - s * w
    out = torch.stack([x0, y0, x1, y1, 2 * cx - x0, 2 * cy - y0, 2 * cx - x1, 2 * cy - y1], dim=-1)
    return out.tolist() if single else out


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
