"""Multi-GPU forms of the hot path (one process per GPU, `torch.distributed`; NCCL over NVLink on B200).

* `sharded_rbbox_overlaps` -- a large rotated-IoU matrix sharded by ROW over the ranks; every rank
  writes its rows of one (m, n) buffer and the shard results are exchanged with ONE in-place
  all-gather (BASELINE.json config C4).
* `scene_merge_nms` -- DOTA scene merge (mmdet/datasets/dota.py:310-336, wwtool.mergebypoly_mp /
  mergebyrec_mp): per-tile per-class NMS on the tiles a rank owns, survivors translated to scene
  coordinates and all-gathered, then the cross-tile merge NMS sharded by CLASS with the per-class
  thresholds of dota.py:321-324 (BASELINE.json config C5).

The compute calls go through the CUDA library (`aidet_b200.ops.functional`); both functions take
an optional compute callable so the host-side logic (partitioning, ragged gathers, ordering) can be
exercised on CPU with the `gloo` backend by the tests, which plug the oracle in.  There is no CPU
compute in this module.
"""
import torch
import torch.distributed as dist

# mmdet/datasets/dota.py:28 (CLASSES) and :321-324 (class-wise merge thresholds)
DOTA_CLASSES = ('harbor', 'ship', 'small-vehicle', 'large-vehicle', 'storage-tank', 'plane', 'soccer-ball-field',
                'bridge', 'baseball-diamond', 'tennis-court', 'helicopter', 'roundabout', 'swimming-pool',
                'ground-track-field', 'basketball-court')
DOTA_OBB_MERGE_THR = {'harbor': 0.1, 'ship': 0.05, 'small-vehicle': 0.15, 'large-vehicle': 0.5, 'storage-tank': 0.35,
                      'plane': 0.2, 'soccer-ball-field': 0.2, 'bridge': 0.45, 'baseball-diamond': 0.2,
                      'tennis-court': 0.1, 'helicopter': 0.1, 'roundabout': 0.15, 'swimming-pool': 0.05,
                      'ground-track-field': 0.4, 'basketball-court': 0.2}
DOTA_HBB_MERGE_THR = {'harbor': 0.4, 'ship': 0.4, 'small-vehicle': 0.4, 'large-vehicle': 0.5, 'storage-tank': 0.1,
                      'plane': 0.25, 'soccer-ball-field': 0.2, 'bridge': 0.5, 'baseball-diamond': 0.15,
                      'tennis-court': 0.2, 'helicopter': 0.2, 'roundabout': 0.15, 'swimming-pool': 0.2,
                      'ground-track-field': 0.15, 'basketball-court': 0.2}


def merge_thresholds(task='obb', classwise=True):
    """(15,) thresholds in DOTA_CLASSES order; classwise=False -> 0.3 everywhere (dota.py:326-331)."""
    table = DOTA_OBB_MERGE_THR if task == 'obb' else DOTA_HBB_MERGE_THR
    return torch.tensor([table[c] if classwise else 0.3 for c in DOTA_CLASSES], dtype=torch.float32)


SOLO = "solo"        # group argument: treat this rank as a world of one (every rank works on its own data, no collective)


def _world(group):
    if isinstance(group, str) and group == SOLO:
        return 1, 0
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rows(m, world, rank):
    """Equal row blocks (the last ones may be short or empty): (rows_per_rank, first row, one past last row)."""
    rows_per = (m + world - 1) // world if m > 0 else 0
    r0 = min(m, rank * rows_per)
    return rows_per, r0, min(m, r0 + rows_per)


def _riou_cuda(a, b, mode, out):
    from .ops import functional as F
    return F.riou_matrix(a, b, mode, out=out)


def sharded_rbbox_overlaps(rbboxes1, rbboxes2, mode='iou', group=None, gather=True, out=None, overlaps_fn=None):
    """Row-sharded (m, n) overlap matrix.

    rbboxes1 (m, 5|8) and rbboxes2 (n, 5|8) are replicated on every rank (100k boxes = 2 MB).  Rank r
    computes rows [r*ceil(m/G), (r+1)*ceil(m/G)) into its slice of a (G*ceil(m/G), n) buffer; with
    gather=True one in-place all-gather makes every rank hold the whole matrix, returned as out[:m].
    With gather=False only the local slice is valid and (rows, (r0, r1)) is returned.
    """
    world, rank = _world(group)
    m, n = rbboxes1.size(0), rbboxes2.size(0)
    rows_per, r0, r1 = shard_rows(m, world, rank)
    fn = overlaps_fn or _riou_cuda
    if out is None:
        out = torch.empty((rows_per * world, n), dtype=torch.float32, device=rbboxes1.device)
    assert out.shape == (rows_per * world, n) and out.is_contiguous()
    mine = out[rank * rows_per:(rank + 1) * rows_per]
    if r1 > r0 and n > 0:
        fn(rbboxes1[r0:r1].contiguous(), rbboxes2, mode, mine[:r1 - r0])
    if not gather:
        return mine[:r1 - r0], (r0, r1)
    if world > 1 and rows_per > 0 and n > 0:
        dist.all_gather_into_tensor(out, mine, group=group)          # in place: rank r's block is already at row r*rows_per
    return out[:m]


class SymmetricMatrix:
    """An (m_pad, n) float32 result buffer allocated in CUDA symmetric memory on every rank of `group`, so each
    rank can store straight into the others' copies (NVLink peer stores).  Build once, reuse across calls."""

    def __init__(self, m, n, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        world, rank = _world(group)
        self.world, self.rank, self.m, self.n = world, rank, m, n
        self.rows_per = shard_rows(m, world, rank)[0]
        self.tensor = symm_mem.empty((self.rows_per * world, n), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        # NVSwitch multicast mapping of the same allocation (0 when the fabric / driver has no NVLS)
        self.mc_ptr = int(getattr(self.handle, 'multicast_ptr', 0) or 0)

    def barrier(self):
        self.handle.barrier()


def sharded_rbbox_overlaps_fused(rbboxes1, rbboxes2, sym, mode='iou', multicast=False):
    """Row-sharded overlap matrix with the all-gather fused into the kernel: rank r computes its row block and
    stores every tile to the same rows of ALL ranks' `sym` buffers (aidet_riou_matrix_multi_f32), so when the
    kernels and the closing barrier are done every rank holds the whole matrix.  No NCCL call on the data path.
    multicast=True (needs `sym.mc_ptr`): every element is stored once to the NVSwitch multicast address and the
    switch replicates it (aidet_riou_matrix_mcast_f32) -- the rank's link carries its block once, not world-1 times.
    """
    from .ops import functional as F
    m, n = rbboxes1.size(0), rbboxes2.size(0)
    assert (m, n) == (sym.m, sym.n)
    rows_per, r0, r1 = shard_rows(m, sym.world, sym.rank)
    sym.barrier()                    # nobody is still reading the previous result
    if r1 > r0 and n > 0 and multicast:
        if not sym.mc_ptr:
            raise RuntimeError("this symmetric-memory allocation has no multicast mapping (no NVLS on this box)")
        F.riou_matrix_mcast(rbboxes1[r0:r1].contiguous(), rbboxes2, sym.mc_ptr + sym.rank * rows_per * n * 4, n, mode)
    elif r1 > r0 and n > 0:
        off = sym.rank * rows_per * n * 4
        # own copy first, then the peers in ring order so the ranks do not all target the same GPU at once
        dst = [sym.ptrs[(sym.rank + q) % sym.world] + off for q in range(sym.world)]
        F.riou_matrix_multi(rbboxes1[r0:r1].contiguous(), rbboxes2, dst, n, mode)
    sym.barrier()                    # every rank's stores have landed
    return sym.tensor[:m]


def gather_ragged(t, group=None):
    """All-gather tensors whose first dimension differs per rank: -> (cat over ranks, counts list)."""
    world, _ = _world(group)
    if world == 1:
        return t, [t.size(0)]
    cnt = torch.tensor([t.size(0)], dtype=torch.int64, device=t.device)
    cnts = torch.empty((world,), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = [int(c) for c in cnts.tolist()]
    mx = max(counts)
    if mx == 0:
        return t, counts
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.size(0)] = t
    buf = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * mx:r * mx + counts[r]] for r in range(world)]), counts


def translate_to_scene(boxes, origins):
    """Tile-frame boxes -> scene frame.  boxes (k, 4|5|8), origins (k, 2) = the (x, y) of each box's tile
    (the `x___y` suffix of the DOTA tile names, tools/dota/dota_demo.py:33)."""
    out = boxes.clone()
    d = boxes.size(1)
    if d == 5:
        out[:, 0] += origins[:, 0]
        out[:, 1] += origins[:, 1]
    else:                                   # 4 (x1,y1,x2,y2) or 8 (x1,y1,...,x4,y4): x at even, y at odd columns
        out[:, 0::2] += origins[:, :1]
        out[:, 1::2] += origins[:, 1:]
    return out


def _nms_cuda(boxes, scores, groups, thr, n_groups):
    from .ops import functional as F
    return F.nms_batched(boxes, scores, groups, thr, n_groups=n_groups, cmp_ge=False, plus_one=False)


def _keep_mask(boxes, scores, groups, thr, n_groups, nms_fn):
    """Batched NMS -> boolean keep mask over ALL boxes.  Boxes whose group id is `n_groups` (not this rank's share) are
    never scanned and never kept.  With the CUDA library nothing touches the host: `keep` is pre-filled with n, the kernels
    overwrite its first n_keep entries, and one indexed store into an (n + 1,) mask does the rest."""
    n = boxes.size(0)
    if nms_fn is not None:                                   # tests: a host callable that returns keep indices
        mask = torch.zeros(n, dtype=torch.bool, device=boxes.device)
        ok = groups < n_groups
        if bool(ok.any()):
            idx = ok.nonzero().flatten()
            mask[idx[nms_fn(boxes[idx].contiguous(), scores[idx].contiguous(), groups[idx].contiguous(), thr, n_groups)]] = True
        return mask
    from .ops import functional as F
    # the kernels write the first n_keep entries of `keep`; the rest keep the fill value n and land in the spare slot
    keep, _ = F.nms_batched(boxes, scores, groups, thr, n_groups=n_groups, cmp_ge=False, plus_one=False, sync=False,
                            keep_fill=n)
    mask = torch.zeros(n + 1, dtype=torch.bool, device=boxes.device)
    mask[keep] = True
    return mask[:n]


def _scene_merge_native(boxes, scores, labels, tile_ids, tile_origins, num_classes, tile_iou_thr, merge_thr, world, rank,
                        group):
    """scene_merge_nms through the library's three scene entry points (include/aidet_b200.h): group ids, the translation
    to the scene frame, both NMS stages and the class-major compaction are device code launched from C++ -- 5 small
    kernels + 2 NMS calls instead of ~30 torch ops whose launch gaps dominated the call.  The only host synchronisation
    is the read of the output count."""
    import ctypes as C

    from . import _lib as L
    from .ops import functional as F
    lib = L.lib()
    dev = boxes.device
    L.require_cuda(boxes, "boxes")
    n, fmt = boxes.shape
    if merge_thr.numel() != num_classes:
        raise ValueError("merge_thr must hold %d thresholds, got %d" % (num_classes, merge_thr.numel()))
    boxes = boxes.contiguous().float()
    scores = scores.to(dev).contiguous().float()
    lab = labels.to(device=dev, dtype=torch.int32).contiguous()
    tid = tile_ids.to(device=dev, dtype=torch.int32).contiguous()
    org = tile_origins.to(device=dev, dtype=torch.float32).contiguous()
    n_tiles = org.size(0)
    merge_thr = merge_thr.contiguous()
    nbytes = lib.aidet_scene_workspace_bytes(n, n_tiles, num_classes, fmt)
    if nbytes == 0:
        raise ValueError("scene_merge_nms: bad sizes n=%d tiles=%d classes=%d box width %d" % (n, n_tiles, num_classes, fmt))
    ws = F._nms_workspace(dev, nbytes + 128)
    wsp = C.c_void_p((ws.data_ptr() + 127) // 128 * 128)
    stream = L.stream_ptr(dev)
    surv = torch.empty(n, dtype=torch.uint8, device=dev)
    sb = torch.empty_like(boxes)
    L.check(lib.aidet_scene_tile_nms_f32(L.dptr(boxes), fmt, L.dptr(scores), L.dptr(lab), L.dptr(tid), L.dptr(org), n, n_tiles,
                                         num_classes, L.dptr(F._scalar_thr(tile_iou_thr, dev)), world, rank, L.dptr(surv),
                                         L.dptr(sb), wsp, nbytes, dev.index, stream), "aidet_scene_tile_nms_f32")
    if world > 1:
        dist.all_reduce(surv, group=group)                   # the ranks' masks are disjoint: the sum is their union
    kept = torch.empty(n, dtype=torch.uint8, device=dev)
    L.check(lib.aidet_scene_merge_nms_f32(L.dptr(sb), fmt, L.dptr(scores), L.dptr(lab), L.dptr(surv), n, n_tiles, num_classes,
                                          L.dptr(merge_thr), world, rank, L.dptr(kept), wsp, nbytes, dev.index, stream),
            "aidet_scene_merge_nms_f32")
    if world > 1:
        dist.all_reduce(kept, group=group)
    out_b = torch.empty_like(sb)
    out_s = torch.empty_like(scores)
    out_l = torch.empty(n, dtype=torch.int32, device=dev)
    out_i = torch.empty(n, dtype=torch.int32, device=dev)
    n_out = torch.empty(1, dtype=torch.int32, device=dev)
    L.check(lib.aidet_scene_compact_f32(L.dptr(sb), fmt, L.dptr(scores), L.dptr(lab), L.dptr(kept), n, num_classes, L.dptr(out_b),
                                        L.dptr(out_s), L.dptr(out_l), L.dptr(out_i), L.dptr(n_out), wsp, nbytes, dev.index,
                                        stream), "aidet_scene_compact_f32")
    k = int(n_out.item())
    return out_b[:k], out_s[:k], out_l[:k].long()


def scene_merge_nms(boxes, scores, labels, tile_ids, tile_origins, num_classes=15, tile_iou_thr=0.5, merge_thr=None,
                    group=None, nms_fn=None):
    """Per-tile NMS + cross-tile merge NMS of one scene, tiles sharded over the ranks.

    boxes (n, 5|8) in TILE coordinates, scores (n,), labels (n,) in [0, num_classes), tile_ids (n,) in
    [0, T); tile_origins (T, 2).  All inputs are replicated (every rank sees every detection -- they are
    KBs).  Stage 1: rank r runs the per-tile NMS of the tiles t with t % G == r in one batched launch (groups =
    tile x class; the other ranks' detections carry an out-of-range group id, which the kernel sorts behind every
    group and never scans).  Stage 2: the survivors are translated to the scene frame and the merge NMS (groups =
    class, thresholds `merge_thr` (num_classes,), default dota.py:324) is sharded by class the same way.  Each stage
    ends with ONE fixed-size all-reduce of the (n,) keep masks -- no ragged gathers, no size exchange, and no host
    synchronisation before the final compaction.  Suppression is `IoU > thr` in both stages (DOTA_devkit keeps
    `ovr <= thresh`).

    Returns (boxes (k, d) scene frame, scores (k,), labels (k,)) -- identical on every rank, ordered by
    (class, tile, original index) so the result does not depend on the world size.
    """
    world, rank = _world(group)
    dev = boxes.device
    n_tiles = tile_origins.size(0)
    if merge_thr is None:
        merge_thr = merge_thresholds('obb')
    merge_thr = merge_thr.to(device=dev, dtype=torch.float32)
    if nms_fn is None and isinstance(tile_iou_thr, (int, float)):
        return _scene_merge_native(boxes, scores, labels, tile_ids, tile_origins, num_classes, float(tile_iou_thr),
                                   merge_thr, world, rank, group)
    labels = labels.long()
    tile_ids = tile_ids.long()

    def exchange(mask):
        if world == 1:
            return mask
        m = mask.to(torch.int32)
        dist.all_reduce(m, group=group)                      # the ranks' masks are disjoint: the sum is their union
        return m > 0

    # ---- stage 1: per-tile, per-class NMS on my tiles
    g1_all = n_tiles * num_classes
    g1 = tile_ids * num_classes + labels
    if world > 1:
        g1 = torch.where((tile_ids % world) == rank, g1, torch.full_like(g1, g1_all))
    g1 = g1.int()
    surv = exchange(_keep_mask(boxes, scores, g1, tile_iou_thr, g1_all, nms_fn))
    # ---- stage 2: cross-tile merge in the scene frame, sharded by class
    sb = translate_to_scene(boxes, tile_origins.to(dev)[tile_ids])
    mine2 = surv if world == 1 else surv & ((labels % world) == rank)
    g2 = torch.where(mine2, labels, torch.full_like(labels, num_classes)).int()
    kept = exchange(_keep_mask(sb, scores, g2, merge_thr, num_classes, nms_fn))
    # class-major output (the reference writes one Task1_<class>.txt per class, dota.py:296-308); the only host
    # synchronisation of the call is this compaction
    kept_all = kept.nonzero().flatten()
    # 8-bit keys when they fit: ONE radix pass instead of the eight of an int64 sort
    lab_k = labels[kept_all]
    order = torch.argsort(lab_k.to(torch.uint8) if num_classes <= 255 else lab_k, stable=True)
    kept_all = kept_all[order]
    return sb[kept_all], scores[kept_all], labels[kept_all]
