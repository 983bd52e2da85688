"""GPU: RPN proposal generation (rpn_get_bboxes: all levels of all images through ONE batched NMS launch) against the
reference's per-image, per-level procedure (mmdet/models/anchor_heads/rpn_head.py:55-108) restated step by step with
CPU torch ops and the oracle NMS (`>` like the CUDA entry, `+1` widths)."""
import numpy as np
import pytest
import torch

from aidet_b200.core import delta2bbox
from aidet_b200.models import rpn_get_bboxes, rpn_get_bboxes_single
from aidet_b200.models.anchor_heads.rpn_head import decode_levels, select_proposals
from oracle import oracle as O

pytestmark = pytest.mark.gpu
STRIDES = (4, 8, 16, 32, 64)


def hbb_anchors(tile, stride, scale=8.0, ratios=(0.5, 1.0, 2.0)):
    """AnchorGenerator.grid_anchors order: cell row, cell column, ratio (mmdet/core/anchor/anchor_generator.py:38-84)."""
    cells = tile // stride
    ys, xs = torch.meshgrid(torch.arange(cells, dtype=torch.float32), torch.arange(cells, dtype=torch.float32), indexing='ij')
    ctr = torch.stack([xs, ys], -1).reshape(-1, 1, 2) * stride + (stride - 1) / 2
    r = torch.tensor(ratios).view(1, -1)
    w, h = scale * stride / r.sqrt(), scale * stride * r.sqrt()
    half = torch.stack([w, h], -1) / 2
    return torch.cat([ctr - half + 0.5, ctr + half - 0.5], -1).reshape(-1, 4)


def make_inputs(n_img, tile, seed):
    g = torch.Generator().manual_seed(seed)
    cls, reg, anc = [], [], []
    for st in STRIDES:
        c = tile // st
        cls.append(torch.randn(n_img, 3, c, c, generator=g) * 2)
        reg.append(torch.randn(n_img, 12, c, c, generator=g) * 0.4)
        anc.append(hbb_anchors(tile, st))
    return cls, reg, anc


def reference_single(cls_scores, bbox_preds, anchors, img_shape, cfg):
    """rpn_head.py:55-108 with CPU torch ops; the NMS is the oracle's (greedy, `+1`, `>`)."""
    mlvl = []
    for s, d, a in zip(cls_scores, bbox_preds, anchors):
        scores = s.permute(1, 2, 0).reshape(-1).sigmoid()
        d = d.permute(1, 2, 0).reshape(-1, 4)
        if cfg['nms_pre'] > 0 and scores.shape[0] > cfg['nms_pre']:
            _, topk = scores.topk(cfg['nms_pre'])
            d, a, scores = d[topk], a[topk], scores[topk]
        p = delta2bbox(a, d, (0, 0, 0, 0), (1, 1, 1, 1), img_shape)
        if cfg['min_bbox_size'] > 0:
            ok = (p[:, 2] - p[:, 0] + 1 >= cfg['min_bbox_size']) & (p[:, 3] - p[:, 1] + 1 >= cfg['min_bbox_size'])
            p, scores = p[ok], scores[ok]
        mlvl.append(torch.cat([p, scores.unsqueeze(-1)], -1))
    return mlvl


def reference_select(mlvl, cfg):
    out = []
    for p in mlvl:
        keep, _ = O.nms(p[:, :4].numpy(), p[:, 4].numpy(), cfg['nms_thr'], cmp_ge=False, plus_one=True)
        out.append(p[torch.from_numpy(keep)][:cfg['nms_post']])
    p = torch.cat(out, 0)
    if cfg['nms_across_levels']:
        keep, _ = O.nms(p[:, :4].numpy(), p[:, 4].numpy(), cfg['nms_thr'], cmp_ge=False, plus_one=True)
        return p[torch.from_numpy(keep)][:cfg['max_num']]
    _, topk = p[:, 4].topk(min(cfg['max_num'], p.shape[0]))
    return p[topk]


def canon(t):
    """rows sorted by (score descending, then x1, y1, x2, y2)"""
    n = t.numpy()
    return torch.from_numpy(n[np.lexsort((n[:, 3], n[:, 2], n[:, 1], n[:, 0], -n[:, 4]))])


CFG = dict(nms_across_levels=False, nms_pre=2000, nms_post=2000, max_num=2000, nms_thr=0.7, min_bbox_size=0)   # configs/dota/*: test_cfg.rpn


def test_decode_matches_cpu_reference(cuda):
    cls, reg, anc = make_inputs(2, 256, seed=1)
    shapes = [(256, 256, 3), (200, 240, 3)]
    props, gids = decode_levels([c.to(cuda) for c in cls], [r.to(cuda) for r in reg], [a.to(cuda) for a in anc], shapes, CFG)
    props, gids = props.cpu(), gids.cpu()
    for i in range(2):
        ref = reference_single([c[i] for c in cls], [r[i] for r in reg], anc, shapes[i], CFG)
        for lvl, p in enumerate(ref):
            got = props[gids == lvl * 2 + i]
            assert got.shape == p.shape
            # same candidates in the same (descending score) order, up to float rounding of sigmoid / exp
            assert (got[:, 4] - p[:, 4]).abs().max() < 1e-6
            assert (got[:, :4] - p[:, :4]).abs().max() < 2e-3
        assert float(props[gids % 2 == i][:, [0, 2]].max()) <= shapes[i][1] - 1
        assert float(props[gids % 2 == i][:, [1, 3]].max()) <= shapes[i][0] - 1


@pytest.mark.parametrize("across,min_size,nms_post,max_num", [(False, 0, 2000, 2000), (False, 6, 300, 500), (True, 0, 1000, 700)])
def test_selection_matches_reference_procedure(cuda, across, min_size, nms_post, max_num):
    """Stage 2 on the device's own decoded proposals: bit-exact against the per-image, per-level loop."""
    cfg = dict(CFG, nms_across_levels=across, min_bbox_size=min_size, nms_post=nms_post, max_num=max_num)
    n_img = 3
    cls, reg, anc = make_inputs(n_img, 256, seed=2)
    shapes = [(256, 256, 3)] * n_img
    props, gids = decode_levels([c.to(cuda) for c in cls], [r.to(cuda) for r in reg], [a.to(cuda) for a in anc], shapes, cfg)
    got = select_proposals(props, gids, n_img, len(STRIDES), cfg)
    pc, gc = props.cpu(), gids.cpu()
    for i in range(n_img):
        mlvl = [pc[gc == lvl * n_img + i] for lvl in range(len(STRIDES))]
        ref = reference_select(mlvl, cfg)
        assert got[i].shape == ref.shape and ref.shape[0] > 50
        if across:
            assert torch.equal(got[i].cpu(), ref)
        else:
            # top-k by score: rows with EQUAL float32 scores may come in any order (Tensor.topk does not define it) and
            # ties at the cut may pick different rows -> compare the score column exactly, the rows as sets above the cut
            a, b = canon(got[i].cpu()), canon(ref)
            assert torch.equal(a[:, 4], b[:, 4])
            above = a[:, 4] > a[-1, 4]
            assert torch.equal(a[above], b[above])


def test_public_entry_points(cuda):
    cls, reg, anc = make_inputs(2, 128, seed=3)
    metas = [dict(img_shape=(128, 128, 3), scale_factor=1.0), dict(img_shape=(120, 100, 3), scale_factor=1.0)]
    dc, dr, da = [c.to(cuda) for c in cls], [r.to(cuda) for r in reg], [a.to(cuda) for a in anc]
    both = rpn_get_bboxes(dc, dr, da, metas, CFG)
    assert len(both) == 2 and all(p.shape[1] == 5 for p in both)
    for i in range(2):         # batching over images does not change an image's proposals
        one = rpn_get_bboxes_single([c[i] for c in dc], [r[i] for r in dr], da, metas[i]['img_shape'], 1.0, CFG)
        s = both[i][:, 4]
        assert bool((s[:-1] >= s[1:]).all())                 # top-k by score (rpn_head.py:103-107)
        assert torch.equal(canon(one.cpu()), canon(both[i].cpu()))
    with pytest.raises(NotImplementedError):
        rpn_get_bboxes(cls, reg, anc, metas, CFG)


def test_merge_aug_proposals(cuda):
    """merge_augs.py:8-43: proposals of two augmented views (one flipped and scaled) mapped back, one NMS, top-k."""
    from aidet_b200.core import bbox_mapping, merge_aug_proposals
    g = torch.Generator().manual_seed(6)
    xy = torch.rand(300, 2, generator=g) * 150
    wh = torch.rand(300, 2, generator=g) * 60 + 8
    base = torch.cat([xy, xy + wh, torch.rand(300, 1, generator=g)], 1)
    view2 = base.clone()
    view2[:, :4] = bbox_mapping(base[:, :4] + torch.randn(300, 4, generator=g), (240, 260, 3), 1.5, True)
    view2[:, 4] = torch.rand(300, generator=g)
    metas = [dict(img_shape=(240, 260, 3), scale_factor=1.0, flip=False), dict(img_shape=(240, 260, 3), scale_factor=1.5, flip=True)]
    cfg = dict(nms_thr=0.7, max_num=100)
    got = merge_aug_proposals([base.to(cuda), view2.to(cuda)], metas, cfg).cpu()
    # restated on the CPU with the oracle NMS (`>`: CUDA tensors take the CUDA comparison, nms_kernel.cu:61)
    from aidet_b200.core import bbox_mapping_back
    rec = view2.clone()
    rec[:, :4] = bbox_mapping_back(view2[:, :4], (240, 260, 3), 1.5, True)
    allp = torch.cat([base, rec], 0)
    keep, _ = O.nms(allp[:, :4].numpy(), allp[:, 4].numpy(), 0.7, cmp_ge=False, plus_one=True)
    ref = allp[torch.from_numpy(keep)]
    ref = ref[ref[:, 4].sort(descending=True)[1][:100]]
    assert got.shape == ref.shape == (100, 5)
    assert torch.allclose(got, ref, atol=1e-4)
