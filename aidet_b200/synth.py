"""Synthetic DOTA-shaped inputs shared by tests/ and bench.py.

All generators are seeded `torch.Generator(device='cpu')` streams and return float32 CPU
tensors; callers move them to the GPU.  Nothing here is part of the compute path.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed))
    return g


def _uniform(g, n, lo, hi):
    return torch.rand(n, generator=g, dtype=torch.float64) * (hi - lo) + lo


def dota_boxes(n, side=1024, seed=0, dense=False):
    """(n,5) theta-OBBs (cx,cy,w,h,theta[rad]) and (n,) scores.

    DOTA-shaped: 70 % of the centres around n/40 cluster centres (sigma 24 px), 30 % uniform;
    long side exp(U(ln 8, ln 256)), aspect U(1,6), theta U(-pi/2, pi/2), scores U(0,1).
    dense=True: every centre inside one 16 px disc, long side exp(U(ln 96, ln 256)), aspect
    U(1,1.5) -- every pair truly intersects, so no early-out can fire (roofline runs).
    """
    g = _gen(seed)
    if dense:
        r = 16.0 * torch.sqrt(_uniform(g, n, 0, 1))
        ph = _uniform(g, n, 0, 2 * math.pi)
        cx = side / 2 + r * torch.cos(ph)
        cy = side / 2 + r * torch.sin(ph)
        long_side = torch.exp(_uniform(g, n, math.log(96), math.log(256)))
        aspect = _uniform(g, n, 1, 1.5)
    else:
        n_clu = max(n // 40, 1)
        centres = _uniform(g, 2 * n_clu, 0, side).view(n_clu, 2)
        which = torch.randint(0, n_clu, (n,), generator=g)
        clustered = _uniform(g, n, 0, 1) < 0.7
        jitter = torch.randn(n, 2, generator=g, dtype=torch.float64) * 24.0
        uni = _uniform(g, 2 * n, 0, side).view(n, 2)
        c = torch.where(clustered[:, None], centres[which] + jitter, uni)
        cx, cy = c[:, 0], c[:, 1]
        long_side = torch.exp(_uniform(g, n, math.log(8), math.log(256)))
        aspect = _uniform(g, n, 1, 6)
    swap = _uniform(g, n, 0, 1) < 0.5
    w = torch.where(swap, long_side, long_side / aspect)
    h = torch.where(swap, long_side / aspect, long_side)
    theta = _uniform(g, n, -math.pi / 2, math.pi / 2)
    scores = _uniform(g, n, 0, 1)
    boxes = torch.stack([cx, cy, w, h, theta], dim=1).float()
    return boxes, scores.float()


def regression_pairs(n, side=1024, seed=0, shift=0.15, scale=0.35, dtheta=0.3):
    """(pred, target) aligned (n,5) theta-OBB pairs as a box-regression loss sees them: targets are DOTA-shaped
    boxes, predictions the same boxes with the centre moved by N(0, shift * sqrt(w h)), sides scaled by
    U(1 - scale, 1 + scale) and the angle turned by N(0, dtheta) -- so nearly every pair overlaps."""
    target, _ = dota_boxes(n, side=side, seed=seed)
    g = _gen(seed + 7919)
    t = target.double()
    size = torch.sqrt(t[:, 2] * t[:, 3])
    d = torch.randn(n, 2, generator=g, dtype=torch.float64) * (shift * size)[:, None]
    sc = _uniform(g, 2 * n, 1 - scale, 1 + scale).view(n, 2)
    dth = torch.randn(n, generator=g, dtype=torch.float64) * dtheta
    pred = torch.stack([t[:, 0] + d[:, 0], t[:, 1] + d[:, 1], t[:, 2] * sc[:, 0], t[:, 3] * sc[:, 1], t[:, 4] + dth], 1)
    return pred.float(), target


def anchor_grid(tile=1024, strides=(4, 8, 16, 32, 64), scale=8.0, ratios=(0.5, 1.0, 2.0), theta=-math.pi / 2):
    """(n,5) theta-OBB anchors of one tile in AnchorGenerator order (level, then cell row, cell column, ratio:
    mmdet/core/anchor/anchor_generator.py:38-84), side = scale * stride, base angle -pi/2 as the OBB codecs use
    (mmdet/core/rbbox/transforms.py:383-390).  1024 tile, P2-P6, 3 ratios -> 261888 anchors."""
    out = []
    for st in strides:
        cells = tile // st
        ys, xs = torch.meshgrid(torch.arange(cells, dtype=torch.float64), torch.arange(cells, dtype=torch.float64),
                                indexing='ij')
        cx = (xs.reshape(-1, 1) + 0.5) * st
        cy = (ys.reshape(-1, 1) + 0.5) * st
        r = torch.tensor(ratios, dtype=torch.float64).view(1, -1)
        w = scale * st / torch.sqrt(r)
        h = scale * st * torch.sqrt(r)
        n = cells * cells
        a = torch.stack([cx.expand(n, len(ratios)), cy.expand(n, len(ratios)), w.expand(n, len(ratios)),
                         h.expand(n, len(ratios)), torch.full((n, len(ratios)), theta, dtype=torch.float64)], dim=2)
        out.append(a.reshape(-1, 5))
    return torch.cat(out).float()


def assign_case(n_boxes, n_gts, side=1024, seed=0, n_ignore=0):
    """(bboxes (n,5), gt (k,5), gt_ignore (q,5), gt_labels (k,)): truths are DOTA-shaped boxes, candidates a mix of
    jittered copies of truths (positives of every quality, exact duplicates included so that a truth's best
    overlap is tied between candidates) and unrelated DOTA-shaped boxes (negatives)."""
    gt, _ = dota_boxes(n_gts, side=side, seed=seed)
    ign, _ = dota_boxes(max(n_ignore, 1), side=side, seed=seed + 1)
    bg, _ = dota_boxes(n_boxes, side=side, seed=seed + 2)
    g = _gen(seed + 104729)
    n_pos = n_boxes // 3
    src = torch.randint(0, n_gts, (n_pos,), generator=g)
    pos, _ = regression_pairs(n_gts, side=side, seed=seed)          # jittered versions of the same truths
    q = _uniform(g, n_pos, 0, 1)[:, None].float()
    boxes = bg.clone()
    boxes[:n_pos] = gt[src] * (1 - q) + pos[src] * q                 # from exact copies (q=0) to loose matches
    n_dup = min(8, n_pos // 2)
    boxes[n_pos:n_pos + n_dup] = boxes[:n_dup]                       # exact duplicates -> tied maxima
    labels = torch.randint(1, 16, (n_gts,), generator=g)
    return boxes, gt, ign[:n_ignore], labels


def thetaobb2pointobb(boxes):
    """(n,5) -> (n,8) in cv2.boxPoints order (mmdet/core/rbbox/transforms.py:45-55), float64 math."""
    b = boxes.double()
    cx, cy, w, h, th = b.unbind(1)
    c, s = torch.cos(th) * 0.5, torch.sin(th) * 0.5
    x0 = cx - s * h - c * w
    y0 = cy + c * h - s * w
    x1 = cx + s * h - c * w
    y1 = cy - c * h - s * w
    return torch.stack([x0, y0, x1, y1, 2 * cx - x0, 2 * cy - y0, 2 * cx - x1, 2 * cy - y1], dim=1).float()


def free_quads(boxes, rel_noise=0.1, seed=0):
    """(n,5) theta-OBBs -> (n,8) free convex quadrilaterals (what a point-OBB head regresses):
    the rectangle's corners jittered by N(0, rel_noise * short side).  Returns (quads, convex mask)."""
    g = _gen(seed)
    p = thetaobb2pointobb(boxes)
    short = torch.minimum(boxes[:, 2], boxes[:, 3])[:, None]
    q = (p + torch.randn(p.shape, generator=g) * rel_noise * short).contiguous()
    v = q.view(-1, 4, 2).double()
    e = torch.roll(v, -1, 1) - v
    cr = e[:, :, 0] * torch.roll(e, -1, 1)[:, :, 1] - e[:, :, 1] * torch.roll(e, -1, 1)[:, :, 0]
    convex = (cr > 0).all(1) | (cr < 0).all(1)
    return q, convex


def multiclass_dets(n=2000, num_classes=15, side=1024, seed=2, dense=False, dim=5):
    """Config C2: (n, (C+1)*5) class-specific theta-OBBs and (n, C+1) softmax scores.

    Every class regresses a jittered copy of the proposal; scores = softmax(3 N(0,1)) over
    C+1 columns (column 0 = background), to be filtered at > 0.05 (rbbox_nms.py:30).
    """
    g = _gen(seed)
    base, _ = dota_boxes(n, side=side, seed=seed + 1000, dense=dense)
    jit = torch.randn(n, num_classes + 1, 5, generator=g) * torch.tensor([2.0, 2.0, 1.0, 1.0, 0.02])
    boxes = base[:, None, :] + jit
    boxes[..., 2:4] = boxes[..., 2:4].clamp(min=2.0)
    scores = torch.softmax(3.0 * torch.randn(n, num_classes + 1, generator=g), dim=1)
    if dim == 8:
        boxes = thetaobb2pointobb(boxes.view(-1, 5)).view(n, num_classes + 1, 8)
    return boxes.reshape(n, -1).contiguous(), scores.contiguous()


def fpn_features(batch=8, channels=256, tile=1024, strides=(4, 8, 16, 32), seed=3):
    """Config C3 features: list of NHWC float32 randn maps (N, tile/s, tile/s, C)."""
    g = _gen(seed)
    return [torch.randn(batch, tile // s, tile // s, channels, generator=g) for s in strides]


def rotated_rois(rois_per_img=512, batch=8, tile=1024, seed=3, num_levels=4, finest_scale=56):
    """Config C3 RoIs: (K,6) [b,cx,cy,w,h,theta] + level ids by the rule of
    mmdet/models/roi_extractors/single_level.py:69-73 (scale = sqrt(w*h))."""
    g = _gen(seed + 77)
    k = rois_per_img * batch
    side = torch.exp(_uniform(g, k, math.log(16), math.log(512)))
    aspect = _uniform(g, k, 1, 4)
    swap = _uniform(g, k, 0, 1) < 0.5
    w = torch.where(swap, side, side / aspect)
    h = torch.where(swap, side / aspect, side)
    theta = _uniform(g, k, -math.pi / 2, math.pi / 2)
    cx = _uniform(g, k, 0, tile)
    cy = _uniform(g, k, 0, tile)
    b = torch.arange(batch, dtype=torch.float64).repeat_interleave(rois_per_img)
    rois = torch.stack([b, cx, cy, w, h, theta], dim=1).float()
    scale = torch.sqrt(rois[:, 3] * rois[:, 4])
    lvl = torch.floor(torch.log2(scale / finest_scale + 1e-6)).clamp(min=0, max=num_levels - 1).int()
    return rois, lvl


def scene_tiles(scene=4000, tile=1024, overlap=200):
    """Config C5 tile origins: stride tile-overlap, last one clamped to scene-tile."""
    step = tile - overlap
    xs = list(range(0, scene - tile, step)) + [scene - tile]
    return [(x, y) for y in xs for x in xs]


def scene_dets(scene=4000, tile=1024, overlap=200, dets_per_tile=2000, num_classes=15, seed=6, dup=3):
    """Config C5: detections of one `scene` x `scene` DOTA image cut into `tile` tiles (`overlap` px overlap).

    Scene-level objects (DOTA-shaped theta-OBBs, class per object) are seen by every tile that contains their
    centre -- objects in the overlap bands therefore appear in 2-4 tiles -- and every sighting produces `dup`
    detections jittered by ~1 px / 0.01 rad (what a detector's surviving proposals look like), so both the
    per-tile NMS and the cross-tile merge have work.  At most `dets_per_tile` detections per tile (top score).

    Returns boxes (n,5) in TILE coordinates, scores (n,), labels (n,) int64, tile_ids (n,) int64, origins (T,2).
    """
    g = _gen(seed)
    origins = torch.tensor(scene_tiles(scene, tile, overlap), dtype=torch.float32)
    n_obj = int(dets_per_tile * origins.size(0) / dup / 1.5)
    obj, _ = dota_boxes(n_obj, side=scene, seed=seed + 500)
    obj_label = torch.randint(0, num_classes, (n_obj,), generator=g)
    obj_score = _uniform(g, n_obj, 0.05, 1.0).float()
    bs, ss, ls, ts = [], [], [], []
    for t, (x0, y0) in enumerate(origins.tolist()):
        inside = ((obj[:, 0] >= x0) & (obj[:, 0] < x0 + tile) & (obj[:, 1] >= y0) & (obj[:, 1] < y0 + tile)).nonzero().flatten()
        if inside.numel() == 0:
            continue
        rep = inside.repeat_interleave(dup)
        b = obj[rep].clone()
        b[:, 0] -= x0
        b[:, 1] -= y0
        b += torch.randn(b.shape, generator=g) * torch.tensor([1.0, 1.0, 1.0, 1.0, 0.01])
        b[:, 2:4] = b[:, 2:4].clamp(min=2.0)
        s = (obj_score[rep] * _uniform(g, rep.numel(), 0.7, 1.0).float()).clamp(max=1.0)
        if s.numel() > dets_per_tile:
            top = torch.topk(s, dets_per_tile).indices.sort().values
            b, s, rep = b[top], s[top], rep[top]
        bs.append(b); ss.append(s); ls.append(obj_label[rep]); ts.append(torch.full((s.numel(),), t, dtype=torch.int64))
    return torch.cat(bs).contiguous(), torch.cat(ss).contiguous(), torch.cat(ls).contiguous(), torch.cat(ts), origins
