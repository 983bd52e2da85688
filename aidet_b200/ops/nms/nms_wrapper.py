"""Call-compatible mirror of mmdet/ops/nms/nms_wrapper.py, served by the sm_100a kernels.

`nms` keeps the reference contract (mmdet/ops/nms/nms_wrapper.py:7-60): tensor or ndarray in,
same container out, `(dets[inds, :], inds)`, inds int64 in ascending ORIGINAL index order
(nms_cpu.cpp:59, nms_kernel.cu:135-138), empty input -> empty long tensor.
`thetaobb_nms` / `pointobb_nms` are the entry points the reference left commented out
(mmdet/core/post_processing/rbbox_nms.py:97); `batched_rnms` replaces the per-class Python
loops (rbbox_nms.py:29-49,83-106) with one launch.

Threshold comparison follows the reference per entry point:
  - CUDA tensor in (or ndarray + device_id)  -> suppress when IoU >  thr (nms_kernel.cu:61)
  - CPU tensor / ndarray without device_id   -> suppress when IoU >= thr (nms_cpu.cpp:56);
    the data is uploaded and the same CUDA kernels run -- there is no CPU implementation.
  - rotated entry points                     -> IoU > thr (DOTA_devkit-style `ovr <= thr` keeps)
"""
import numpy as np
import torch

from .. import functional as F


def _to_device(dets, device_id):
    if isinstance(dets, torch.Tensor):
        return False, dets, dets.is_cuda
    if isinstance(dets, np.ndarray):
        return True, torch.from_numpy(dets), device_id is not None
    raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))


def _cuda_device(dets_th, device_id):
    if dets_th.is_cuda:
        return dets_th.device
    if not torch.cuda.is_available():
        raise RuntimeError('aidet_b200 needs a CUDA device: there is no CPU NMS in this build')
    return torch.device('cuda', torch.cuda.current_device() if device_id is None else device_id)


def _run(dets, iou_thr, device_id, ncol, fmt_name, cuda_semantics, plus_one):
    is_numpy, dets_th, wants_cuda = _to_device(dets, device_id)
    if dets_th.dim() != 2 or dets_th.size(1) != ncol:
        raise ValueError('{} expects dets of shape (N, {}), got {}'.format(fmt_name, ncol, tuple(dets_th.shape)))
    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    else:
        dev = _cuda_device(dets_th, device_id)
        d = dets_th.to(device=dev, dtype=torch.float32)
        cmp_ge = (not wants_cuda) and (not cuda_semantics)
        inds = F.nms_batched(d[:, :ncol - 1], d[:, ncol - 1], None, float(iou_thr), cmp_ge=cmp_ge, plus_one=plus_one)
        inds = inds.to(dets_th.device)
    if is_numpy:
        inds = inds.cpu().numpy()
    return dets[inds, :], inds


def nms(dets, iou_thr, device_id=None):
    """Axis-aligned NMS, dets (N,5) [x1,y1,x2,y2,score], legacy +1 areas.

    Example (mmdet/ops/nms/nms_wrapper.py:25-34):
        >>> dets = np.array([[49.1, 32.4, 51.0, 35.9, 0.9],
        >>>                  [49.3, 32.9, 51.0, 35.3, 0.9],
        >>>                  [49.2, 31.8, 51.0, 35.4, 0.5],
        >>>                  [35.1, 11.5, 39.1, 15.7, 0.5],
        >>>                  [35.6, 11.8, 39.3, 14.2, 0.5],
        >>>                  [35.3, 11.5, 39.9, 14.5, 0.4],
        >>>                  [35.2, 11.7, 39.7, 15.7, 0.3]], dtype=np.float32)
        >>> suppressed, inds = nms(dets, 0.7)
        >>> assert len(inds) == len(suppressed) == 3
    """
    return _run(dets, iou_thr, device_id, 5, 'nms', cuda_semantics=False, plus_one=True)


def thetaobb_nms(dets, iou_thr, device_id=None):
    """Rotated NMS, dets (N,6) [cx,cy,w,h,theta(rad),score] -> (dets[inds,:], inds)."""
    return _run(dets, iou_thr, device_id, 6, 'thetaobb_nms', cuda_semantics=True, plus_one=False)


def pointobb_nms(dets, iou_thr, device_id=None):
    """Polygon NMS, dets (N,9) [x1,y1,...,x4,y4,score] -> (dets[inds,:], inds)."""
    return _run(dets, iou_thr, device_id, 9, 'pointobb_nms', cuda_semantics=True, plus_one=False)


def batched_rnms(rboxes, scores, group_ids, iou_thr, n_groups=None):
    """One-launch NMS over groups (image x class).

    rboxes (N,5|8) or HBB (N,4, no +1), scores (N,), group_ids (N,) int in [0, n_groups),
    iou_thr float or per-group (n_groups,) thresholds (mmdet/datasets/dota.py:324).
    Returns keep (K,) int64, ascending original index.
    """
    return F.nms_batched(rboxes, scores, group_ids, iou_thr, n_groups=n_groups, cmp_ge=False, plus_one=False)


def soft_nms(dets, iou_thr, method='linear', sigma=0.5, min_score=1e-3):
    """Soft-NMS with the reference's contract (mmdet/ops/nms/nms_wrapper.py:63-118): tensor or ndarray in, same
    container out, `(new_dets (k,5) with decayed scores, inds (k,))` in selection order.

    The reference has only a CPU kernel (nms_cpu.cpp:70-201) and round-trips CUDA tensors through the host
    (nms_wrapper.py:92-94,110-114); here the same loop runs on the device (aidet_soft_nms_f32).  CPU tensors and
    ndarrays are uploaded -- there is no CPU implementation.

    Example (nms_wrapper.py:81-90):
        >>> dets = np.array([[4., 3., 5., 3., 0.9],
        >>>                  [4., 3., 5., 4., 0.9],
        >>>                  [3., 1., 3., 1., 0.5],
        >>>                  [3., 1., 3., 1., 0.5],
        >>>                  [3., 1., 3., 1., 0.4],
        >>>                  [3., 1., 3., 1., 0.0]], dtype=np.float32)
        >>> new_dets, inds = soft_nms(dets, 0.7, sigma=0.5)
        >>> assert len(inds) == len(new_dets) == 3
    """
    is_numpy, dets_th, _ = _to_device(dets, None)
    method_codes = {'linear': 1, 'gaussian': 2}
    if method not in method_codes:
        raise ValueError('Invalid method for SoftNMS: {}'.format(method))
    if dets_th.dim() != 2 or dets_th.size(1) != 5:
        raise ValueError('soft_nms expects dets of shape (N, 5), got {}'.format(tuple(dets_th.shape)))
    if dets_th.shape[0] == 0:
        new_dets, inds = dets_th.new_zeros((0, 5)), dets_th.new_zeros(0, dtype=torch.long)
    else:
        dev = _cuda_device(dets_th, None)
        rows, _ = F.soft_nms_batched(dets_th.to(device=dev, dtype=torch.float32), None, float(iou_thr),
                                     method_codes[method], float(sigma), float(min_score))
        new_dets = rows[:, :5].to(device=dets_th.device, dtype=dets_th.dtype)
        inds = rows[:, 5].to(device=dets_th.device, dtype=torch.long)
    if is_numpy:
        return new_dets.numpy().astype(dets.dtype), inds.numpy().astype(np.int64)
    return new_dets, inds
