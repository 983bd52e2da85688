"""`delta2bbox` / `bbox2delta` (mmdet/core/bbox/transforms.py:6-113): the axis-aligned box codec the RPN stage applies
before its NMS.  Tensor ops only (this is the plumbing around the batched NMS launch, not a kernel)."""
import math

import numpy as np
import torch


def bbox2delta(proposals, gt, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """mmdet/core/bbox/transforms.py:6-31 (legacy `+1` widths)."""
    assert proposals.size() == gt.size()
    proposals, gt = proposals.float(), gt.float()
    px = (proposals[..., 0] + proposals[..., 2]) * 0.5
    py = (proposals[..., 1] + proposals[..., 3]) * 0.5
    pw = proposals[..., 2] - proposals[..., 0] + 1.0
    ph = proposals[..., 3] - proposals[..., 1] + 1.0
    gx = (gt[..., 0] + gt[..., 2]) * 0.5
    gy = (gt[..., 1] + gt[..., 3]) * 0.5
    gw = gt[..., 2] - gt[..., 0] + 1.0
    gh = gt[..., 3] - gt[..., 1] + 1.0
    deltas = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph)], dim=-1)
    return (deltas - deltas.new_tensor(means)) / deltas.new_tensor(stds)


def delta2bbox(rois, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    """mmdet/core/bbox/transforms.py:34-113.  rois (..., 4), deltas (..., 4) -> boxes (..., 4) <x1, y1, x2, y2>.

    `max_shape` is (H, W) as in the reference, or a pair of tensors broadcastable against the leading dimensions (one
    bound per image when a whole batch is decoded at once).

    >>> rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    >>> deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]])
    >>> delta2bbox(rois, deltas, max_shape=(32, 32))      # the reference's doctest (:57-71)
    tensor([[0.0000, 0.0000, 1.0000, 1.0000],
            [0.2817, 0.2817, 4.7183, 4.7183],
            [0.0000, 0.6321, 7.3891, 0.3679],
            [5.8967, 2.9251, 5.5033, 3.2749]])
    """
    d = deltas * deltas.new_tensor(stds) + deltas.new_tensor(means)
    max_ratio = abs(math.log(wh_ratio_clip))
    dx, dy = d[..., 0], d[..., 1]
    dw = d[..., 2].clamp(min=-max_ratio, max=max_ratio)
    dh = d[..., 3].clamp(min=-max_ratio, max=max_ratio)
    px = (rois[..., 0] + rois[..., 2]) * 0.5
    py = (rois[..., 1] + rois[..., 3]) * 0.5
    pw = rois[..., 2] - rois[..., 0] + 1.0
    ph = rois[..., 3] - rois[..., 1] + 1.0
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = torch.addcmul(px, pw, dx), torch.addcmul(py, ph, dy)
    x1, y1 = gx - gw * 0.5 + 0.5, gy - gh * 0.5 + 0.5
    x2, y2 = gx + gw * 0.5 - 0.5, gy + gh * 0.5 - 0.5
    if max_shape is not None:
        hmax, wmax = max_shape[0] - 1, max_shape[1] - 1
        if torch.is_tensor(hmax):
            zero = torch.zeros((), dtype=x1.dtype, device=x1.device)
            x1, x2 = torch.min(torch.max(x1, zero), wmax), torch.min(torch.max(x2, zero), wmax)
            y1, y2 = torch.min(torch.max(y1, zero), hmax), torch.min(torch.max(y2, zero), hmax)
        else:
            x1, x2 = x1.clamp(min=0, max=wmax), x2.clamp(min=0, max=wmax)
            y1, y2 = y1.clamp(min=0, max=hmax), y2.clamp(min=0, max=hmax)
    return torch.stack([x1, y1, x2, y2], dim=-1)


def bbox_flip(bboxes, img_shape):
    """Flip boxes horizontally (mmdet/core/bbox/transforms.py:113-131, tensor branch): bboxes (n, 4k), img_shape (h, w, ..)."""
    assert bboxes.shape[-1] % 4 == 0
    flipped = bboxes.clone()
    flipped[:, 0::4] = img_shape[1] - bboxes[:, 2::4] - 1
    flipped[:, 2::4] = img_shape[1] - bboxes[:, 0::4] - 1
    return flipped


def bbox_mapping(bboxes, img_shape, scale_factor, flip):
    """Original image scale -> testing scale (transforms.py:134-139)."""
    new_bboxes = bboxes * scale_factor
    return bbox_flip(new_bboxes, img_shape) if flip else new_bboxes


def bbox_mapping_back(bboxes, img_shape, scale_factor, flip):
    """Testing scale -> original image scale (transforms.py:142-146)."""
    new_bboxes = bbox_flip(bboxes, img_shape) if flip else bboxes
    return new_bboxes / scale_factor


def bbox2roi(bbox_list):
    """List of per-image (n_i, >=4) boxes -> (n, 5) [batch_ind, x1, y1, x2, y2] RoIs (transforms.py:149-168): the input
    format of RoIAlign / SingleRoIExtractor."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def rbbox2roi(rbbox_list):
    """The same for theta-OBBs: per-image (n_i, >=5) <cx, cy, w, h, theta> -> (n, 6) [batch_ind, cx, cy, w, h, theta], the
    input format of RoIAlignRotated."""
    rois_list = []
    for img_id, rb in enumerate(rbbox_list):
        if rb.size(0) > 0:
            rois_list.append(torch.cat([rb.new_full((rb.size(0), 1), img_id), rb[:, :5]], dim=-1))
        else:
            rois_list.append(rb.new_zeros((0, 6)))
    return torch.cat(rois_list, 0)


def roi2bbox(rois):
    """(n, 1 + d) RoIs -> list of per-image (n_i, d) boxes (transforms.py:171-178)."""
    bbox_list = []
    for img_id in torch.unique(rois[:, 0].cpu(), sorted=True):
        bbox_list.append(rois[rois[:, 0] == img_id.item(), 1:])
    return bbox_list


def bbox2result(bboxes, labels, num_classes):
    """(n, 5) detections + (n,) labels -> list of per-class ndarrays (transforms.py:181-199)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)]
    bboxes = bboxes.cpu().numpy()
    labels = labels.cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes - 1)]
