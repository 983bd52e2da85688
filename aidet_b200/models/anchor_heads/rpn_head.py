"""Proposal generation of the RPN head: `RPNHead.get_bboxes_single` (mmdet/models/anchor_heads/rpn_head.py:55-108) and
the per-image loop around it (`AnchorHead.get_bboxes`, mmdet/models/anchor_heads/anchor_head.py:254-278).

The reference runs, per image and per FPN level, top-k -> delta2bbox -> size filter -> `nms` (one sort + mask kernel +
D2H mask copy + host scan each, mmdet/ops/nms/src/nms_kernel.cu:71-139) -> `[:nms_post]`: 5 levels x N images NMS calls
with a host round trip each.  Here the levels of ALL images go through ONE batched launch (`aidet_nms_batched_f32`,
group = level x image, HBB `+1`, suppression on `>` like the CUDA entry); everything else is the same tensor ops in the
same order, so the selected proposals are the reference's.
"""
import torch

from ...core.bbox.transforms import delta2bbox
from ...ops import functional as F


def _cfg(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def decode_levels(cls_scores, bbox_preds, mlvl_anchors, img_shapes, cfg, use_sigmoid_cls=True,
                  target_means=(.0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0)):
    """Stage 1 (rpn_head.py:64-92 for every image at once): scores, top-`nms_pre` per (level, image), delta2bbox with the
    image's own bounds, `min_bbox_size` filter.

    cls_scores[l] (N, A[*2], H, W), bbox_preds[l] (N, 4A, H, W), mlvl_anchors[l] (A*H*W, 4), img_shapes: N x (h, w, ...).
    Returns flat (P, 5) <x1, y1, x2, y2, score> and (P,) group ids `level * N + image`, blocks in ascending group order,
    inside a block the reference's order (descending score when the top-k ran, anchor order otherwise).
    """
    n_img = cls_scores[0].size(0)
    dev = cls_scores[0].device
    hs = torch.tensor([float(s[0]) for s in img_shapes], device=dev).view(n_img, 1)
    ws = torch.tensor([float(s[1]) for s in img_shapes], device=dev).view(n_img, 1)
    nms_pre, min_size = _cfg(cfg, 'nms_pre', -1), _cfg(cfg, 'min_bbox_size', 0)
    props, gids = [], []
    for lvl in range(len(cls_scores)):
        s, d = cls_scores[lvl], bbox_preds[lvl]
        assert s.size()[-2:] == d.size()[-2:]
        s = s.permute(0, 2, 3, 1)
        if use_sigmoid_cls:
            scores = s.reshape(n_img, -1).sigmoid()
        else:
            scores = s.reshape(n_img, -1, 2).softmax(dim=2)[:, :, 1]
        d = d.permute(0, 2, 3, 1).reshape(n_img, -1, 4)
        anchors = mlvl_anchors[lvl].unsqueeze(0).expand(n_img, -1, -1)
        if nms_pre > 0 and scores.size(1) > nms_pre:
            scores, topk = scores.topk(nms_pre, dim=1)
            idx = topk.unsqueeze(-1).expand(-1, -1, 4)
            d, anchors = d.gather(1, idx), anchors.gather(1, idx)
        boxes = delta2bbox(anchors, d, target_means, target_stds, (hs, ws))
        gid = torch.arange(n_img, device=dev).view(n_img, 1).expand(-1, scores.size(1)) + lvl * n_img
        p = torch.cat([boxes, scores.unsqueeze(-1)], dim=-1).reshape(-1, 5)
        gid = gid.reshape(-1)
        if min_size > 0:
            ok = ((p[:, 2] - p[:, 0] + 1 >= min_size) & (p[:, 3] - p[:, 1] + 1 >= min_size))
            p, gid = p[ok], gid[ok]
        props.append(p)
        gids.append(gid)
    return torch.cat(props, 0), torch.cat(gids, 0)


def select_proposals(props, gids, n_img, n_levels, cfg):
    """Stage 2 (rpn_head.py:94-108): per-(level, image) NMS -- one launch for all of them --, `[:nms_post]`, then per
    image either NMS across levels + `[:max_num]` or top-`max_num` by score.  -> list of N (k_i, 5) tensors.

    The ragged per-image results are cut out of ONE device-side sort; the host reads the per-image counts once."""
    nms_thr, nms_post, max_num = _cfg(cfg, 'nms_thr'), _cfg(cfg, 'nms_post'), _cfg(cfg, 'max_num')
    n_groups = n_img * n_levels
    keep = F.nms_batched(props[:, :4], props[:, 4], gids.int(), nms_thr, n_groups=n_groups, cmp_ge=False, plus_one=True)
    kg = gids[keep]                                            # ascending: group blocks are contiguous in `keep`
    cnt = torch.bincount(kg, minlength=n_groups)
    rank = torch.arange(keep.numel(), device=keep.device) - (torch.cumsum(cnt, 0) - cnt)[kg]
    sel = rank < nms_post                                      # proposals[:cfg.nms_post] of every (level, image)
    img = kg % n_img
    p = props[keep]
    per_img = cnt.clamp(max=nms_post).view(n_levels, n_img).sum(0)
    inf = torch.full((), float('inf'), dtype=torch.float64, device=p.device)
    if _cfg(cfg, 'nms_across_levels', False):
        # level-major order inside an image (= torch.cat(mlvl_proposals)), images back to back, dropped rows last
        key = torch.where(sel, img.double() * keep.numel() + torch.arange(keep.numel(), device=p.device).double(), inf)
        n_valid = int(per_img.sum().item())
        flat = p[torch.argsort(key)[:n_valid]]
        g2 = torch.repeat_interleave(torch.arange(n_img, device=p.device), per_img)
        k2 = F.nms_batched(flat[:, :4], flat[:, 4], g2.int(), nms_thr, n_groups=n_img, cmp_ge=False, plus_one=True)
        c2 = torch.bincount(g2[k2], minlength=n_img).tolist()
        out, off = [], 0
        for c in c2:
            out.append(flat[k2[off:off + min(c, max_num)]])
            off += c
        return out
    # per image: descending score (scores lie in [0, 1]); dropped rows sort to the end
    key = torch.where(sel, img.double() * 2 + (1 - p[:, 4].double()), inf)
    order = torch.argsort(key)
    out, off = [], 0
    for c in per_img.tolist():
        out.append(p[order[off:off + min(c, max_num)]])
        off += c
    return out


def rpn_get_bboxes(cls_scores, bbox_preds, mlvl_anchors, img_metas, cfg, rescale=False, use_sigmoid_cls=True,
                   target_means=(.0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0)):
    """`AnchorHead.get_bboxes` for the RPN head (anchor_head.py:254-278 + rpn_head.py:55-108): a list with one (k, 5)
    proposal tensor per image.  `rescale` is accepted and ignored exactly as rpn_head.py:61 ignores it."""
    assert len(cls_scores) == len(bbox_preds) == len(mlvl_anchors)
    if not cls_scores[0].is_cuda:
        raise NotImplementedError('rpn_get_bboxes has no CPU implementation')
    shapes = [m['img_shape'] for m in img_metas]
    props, gids = decode_levels(cls_scores, bbox_preds, mlvl_anchors, shapes, cfg, use_sigmoid_cls, target_means, target_stds)
    return select_proposals(props, gids, len(img_metas), len(cls_scores), cfg)


def rpn_get_bboxes_single(cls_scores, bbox_preds, mlvl_anchors, img_shape, scale_factor, cfg, rescale=False,
                          use_sigmoid_cls=True, target_means=(.0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0)):
    """`RPNHead.get_bboxes_single` (rpn_head.py:55-108): cls_scores[l] (A, H, W), bbox_preds[l] (4A, H, W) of ONE image."""
    metas = [dict(img_shape=img_shape, scale_factor=scale_factor)]
    return rpn_get_bboxes([s.unsqueeze(0) for s in cls_scores], [d.unsqueeze(0) for d in bbox_preds], mlvl_anchors, metas,
                          cfg, rescale, use_sigmoid_cls, target_means, target_stds)[0]
