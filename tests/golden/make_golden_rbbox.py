"""Generates tests/golden/golden_rbbox_v1.npz by running the REFERENCE's own OBB codecs and target builder
(/root/reference/mmdet/core/rbbox/transforms.py and rbbox_target.py, loaded file by file -- `import mmdet` itself
fails here for want of mmcv / pycocotools) on the CPU.  Run once in the dev container; the .npz is committed and
travels to the GPU box, /root/reference does not.

Shims the loader needs, none of which touches the functions' arithmetic:
  * `pycocotools.mask`, `mmcv`: empty stub modules (imported at the top of the files, used only by functions that are
    out of scope here: maskobb2thetaobb, tensor2imgs);
  * `np.int0`: removed in NumPy 2.0; it was an alias of `np.intp` (truncation towards zero when casting floats),
    restored as exactly that;
  * `torch.addcmul(x, 1, a, b)`: the positional-`value` overload the reference's delta decoders use was removed from
    torch; wrapped to `torch.addcmul(x, a, b, value=1)` for the duration of the run.
cv2 here is 4.13; the angle convention of `cv2.minAreaRect` has changed between cv2 releases, so the goldens pin "what
the reference computes with the cv2 of this image" -- aidet_b200 calls the same cv2 entry point and follows suit.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
if not hasattr(np, "int0"):
    np.int0 = np.intp
for stub in ("pycocotools", "pycocotools.mask", "mmcv"):
    sys.modules.setdefault(stub, types.ModuleType(stub))
sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
for name, path in (("mmdet", "mmdet"), ("mmdet.core", "mmdet/core"), ("mmdet.core.rbbox", "mmdet/core/rbbox"),
                   ("mmdet.core.utils", "mmdet/core/utils")):
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REF, path)]      # a bare namespace: the package's own __init__ is NOT executed
    sys.modules[name] = mod
sys.modules["mmdet.core.utils"].multi_apply = importlib.import_module("mmdet.core.utils.misc").multi_apply
T = importlib.import_module("mmdet.core.rbbox.transforms")
rbbox_target = importlib.import_module("mmdet.core.rbbox.rbbox_target").rbbox_target

_addcmul = torch.addcmul


def _addcmul_compat(inp, *args, **kw):
    if len(args) == 3 and not torch.is_tensor(args[0]):
        return _addcmul(inp, args[1], args[2], value=args[0])
    return _addcmul(inp, *args, **kw)


torch.addcmul = _addcmul_compat

rng = np.random.default_rng(11)
out = {}


def theta_boxes(n, side=800.0):
    c = rng.uniform(60, side - 60, (n, 2))
    long_side = np.exp(rng.uniform(np.log(12), np.log(300), n))
    aspect = rng.uniform(1, 6, n)
    th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    return np.stack([c[:, 0], c[:, 1], long_side, long_side / aspect, th], 1)


# ---- list-in / list-out converters, one box at a time (transforms.py:30-163)
tb = theta_boxes(64)
# axis-parallel cases + one generic angle.  (An exactly diagonal box is left out on purpose: after the integer
# truncation of hobb2pointobb its four corner orders TIE in pointobb_best_point_sort, and the reference resolves the
# tie with np.argsort's default kind, which NumPy 2 dispatches to an unstable SIMD sort on AVX-512 hosts -- the winner is
# platform dependent.  aidet_b200 resolves ties to the smallest shift; see tests/test_rbbox_golden.py.)
tb[:4, 4] = [0.0, np.pi / 2, -np.pi / 2, 0.3]
pts = np.array([T.thetaobb2pointobb(b.tolist()) for b in tb])
out["theta_in"] = tb
out["theta2point"] = pts
out["point2theta"] = np.array([T.pointobb2thetaobb(p.tolist()) for p in pts])
out["point2bbox"] = np.array([T.pointobb2bbox(p.tolist()) for p in pts])
out["extreme_sort"] = np.array([T.pointobb_extreme_sort(p.tolist()) for p in pts])
out["best_point_sort"] = np.array([T.pointobb_best_point_sort(p.tolist()) for p in pts])
out["theta2hobb_best"] = np.array([T.thetaobb2hobb(b.tolist(), T.pointobb_best_point_sort) for b in tb])
out["theta2hobb_extreme"] = np.array([T.thetaobb2hobb(b.tolist(), T.pointobb_extreme_sort) for b in tb])
hb = out["theta2hobb_best"].copy()
out["hobb2point"] = np.array([T.hobb2pointobb(h.tolist()) for h in hb])
# ties of the extreme sort: two top points with equal y (transforms.py:105-109)
tie = np.array([[10., 5., 30., 5., 30., 20., 10., 20.], [30., 5., 30., 20., 10., 20., 10., 5.], [7., 9., 1., 3., 7., 3., 9., 9.]])
out["tie_in"] = tie
out["tie_extreme_sort"] = np.array([T.pointobb_extreme_sort(p.tolist()) for p in tie])
out["tie_best_point_sort"] = np.array([T.pointobb_best_point_sort(p.tolist()) for p in tie])

# ---- flips and test-time-augmentation mappings (transforms.py:191-275, 398-409, 507-519, 602-612)
img_shape = (768, 1024, 3)
out["img_shape"] = np.array(img_shape)
out["thetaobb_flip"] = T.thetaobb_flip(tb.copy(), img_shape)
out["thetaobb_flip_1d"] = T.thetaobb_flip(tb[0].copy(), img_shape)
out["pointobb_flip"] = T.pointobb_flip(pts.copy(), img_shape)
out["pointobb_flip_1d"] = T.pointobb_flip(pts[5].copy(), img_shape)
out["hobb_flip"] = T.hobb_flip(hb.copy(), img_shape)
out["hobb_flip_1d"] = T.hobb_flip(hb[3].copy(), img_shape)
for flip in (False, True):
    tag = "_flip" if flip else ""
    out["thetaobb_mapping" + tag] = T.thetaobb_mapping(tb.copy(), img_shape, 1.5, flip)
    out["thetaobb_mapping_back" + tag] = T.thetaobb_mapping_back(tb.copy(), img_shape, 1.5, flip)
    out["pointobb_mapping" + tag] = T.pointobb_mapping(pts.copy(), img_shape, 0.75, flip)
    out["pointobb_mapping_back" + tag] = T.pointobb_mapping_back(pts.copy(), img_shape, 0.75, flip)
    out["hobb_mapping" + tag] = T.hobb_mapping(hb.copy(), img_shape, 1.25, flip)
    out["hobb_mapping_back" + tag] = T.hobb_mapping_back(hb.copy(), img_shape, 1.25, flip)

# ---- rescale (in place, tensors; transforms.py:280-319) and the delta codecs (transforms.py:321-600)
multi = torch.from_numpy(np.concatenate([tb[:8], tb[8:16]], 1)).float()              # (8, 10): two classes per row
out["thetaobb_rescale_in"] = multi.numpy().copy()
out["thetaobb_rescale"] = T.thetaobb_rescale(multi.clone(), 2.5).numpy()
out["thetaobb_rescale_rev"] = T.thetaobb_rescale(multi.clone(), 2.5, reverse_flag=True).numpy()
half = np.maximum(tb[:, 2:4] * 0.45, 6.0)                                              # proposals stay at least 6 px wide
prop = np.concatenate([tb[:, :2] - half, tb[:, :2] + half], 1) + rng.uniform(-2.5, 2.5, (64, 4))
prop_t = torch.from_numpy(prop).float()
means5, stds5 = [0.0, 0.0, 0.0, 0.0, 0.0], [0.1, 0.1, 0.2, 0.2, 0.1]
means8, stds8 = [0.0] * 8, [0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 0.2, 0.2]
out["prop"] = prop_t.numpy()
out["stds5"], out["stds8"] = np.array(stds5), np.array(stds8)
d5 = T.thetaobb2delta(prop_t, torch.from_numpy(tb).float(), means5, stds5)
out["thetaobb2delta"] = d5.numpy()
out["delta2thetaobb"] = T.delta2thetaobb(prop_t, d5.clone(), means5, stds5).numpy()
d5c = torch.cat([d5, d5.flip(0) * 0.5, d5 * 40.0], 1)                                  # 3 classes per row; the last one clamps dw / dh
out["delta5_multi"] = d5c.numpy()
out["delta2thetaobb_multi"] = T.delta2thetaobb(prop_t, d5c.clone(), means5, stds5).numpy()
d8 = T.pointobb2delta(prop_t, torch.from_numpy(pts).float(), means8, stds8)
out["pointobb2delta"] = d8.numpy()
out["delta2pointobb"] = T.delta2pointobb(prop_t, d8.clone(), means8, stds8).numpy()
d8c = torch.cat([d8, d8.flip(0)], 1)
out["delta8_multi"] = d8c.numpy()
out["delta2pointobb_multi"] = T.delta2pointobb(prop_t, d8c.clone(), means8, stds8).numpy()
dh = T.hobb2delta(prop_t, torch.from_numpy(hb).float(), means5, stds5)
out["hobb_in"] = hb
out["hobb2delta"] = dh.numpy()
out["delta2hobb"] = T.delta2hobb(prop_t, dh.clone(), means5, stds5).numpy()
dhc = torch.cat([dh, dh * 30.0], 1)
out["deltah_multi"] = dhc.numpy()
out["delta2hobb_multi"] = T.delta2hobb(prop_t, dhc.clone(), means5, stds5).numpy()

# ---- rbbox_target (rbbox_target.py:8-88): two images, one of them without negatives
class _Cfg:
    def __init__(self, encode):
        self.encode = encode


gt_sets = {"thetaobb": tb, "pointobb": pts, "hobb": hb}
for encode, dim in (("thetaobb", 5), ("pointobb", 8), ("hobb", 5)):
    gts = torch.from_numpy(gt_sets[encode]).float()
    pos1, neg1 = prop_t[:9], prop_t[20:31]
    pos2, neg2 = prop_t[40:45], prop_t[:0]
    inds1 = torch.tensor([3, 3, 0, 7, 1, 2, 2, 5, 6])
    inds2 = torch.tensor([1, 0, 0, 2, 4])
    labels1 = torch.tensor([4, 9, 1, 15, 2, 7, 7, 3])
    labels2 = torch.tensor([5, 6, 11, 2, 8])
    means, stds = (means8, stds8) if dim == 8 else (means5, stds5)
    res = rbbox_target([pos1, pos2], [neg1, neg2], [inds1, inds2], [gts[:8], gts[30:35]], [labels1, labels2],
                       _Cfg(encode), means, stds, out_dim_reg=dim)
    for nm, v in zip(("labels", "label_weights", "targets", "weights"), res):
        out["target_%s_%s" % (encode, nm)] = v.numpy()
    res = rbbox_target([pos1, pos2], [neg1, neg2], [inds1, inds2], [gts[:8], gts[30:35]], [labels1, labels2],
                       _Cfg(encode), means, stds, out_dim_reg=dim, concat=False)
    out["target_%s_split_sizes" % encode] = np.array([r.shape[0] for r in res[0]])
out["target_inds1"], out["target_inds2"] = inds1.numpy(), inds2.numpy()
out["target_labels1"], out["target_labels2"] = labels1.numpy(), labels2.numpy()

torch.addcmul = _addcmul
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rbbox_v1.npz"), **out)
print("wrote golden_rbbox_v1.npz:", len(out), "arrays")
