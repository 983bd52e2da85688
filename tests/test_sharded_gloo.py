"""CPU, world_size 2, gloo: the host-side logic of aidet_b200.sharded (row partition + in-place all-gather,
tile / class sharding, ragged gathers, world-size-independent ordering).  The compute callable is the
float64 oracle here (tests may use it); on the GPU the same functions call the CUDA library."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_overlaps(a, b, mode, out):
    from oracle import oracle as O
    out.copy_(torch.from_numpy(O.riou_matrix(a.numpy(), b.numpy(), mode=mode).astype(np.float32)))
    return out


def _oracle_nms(boxes, scores, groups, thr, n_groups):
    from oracle import oracle as O
    thr = thr.numpy() if isinstance(thr, torch.Tensor) else thr
    keep, _ = O.nms(boxes.numpy(), scores.numpy(), thr, groups=groups.numpy(), cmp_ge=False, plus_one=False)
    return torch.from_numpy(keep)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aidet_b200 import sharded, synth
        res = {}
        # ---- row-sharded IoU, ragged: 37 rows over 2 ranks (19 + 18), and fewer rows than ranks
        a, _ = synth.dota_boxes(37, side=200, seed=1)
        b, _ = synth.dota_boxes(53, side=200, seed=2)
        full = sharded.sharded_rbbox_overlaps(a, b, overlaps_fn=_oracle_overlaps)
        res["iou"] = full.clone()
        part, (r0, r1) = sharded.sharded_rbbox_overlaps(a, b, gather=False, overlaps_fn=_oracle_overlaps)
        res["iou_part"] = (part.clone(), r0, r1)
        one = sharded.sharded_rbbox_overlaps(a[:1], b, overlaps_fn=_oracle_overlaps)
        res["iou_one"] = one.clone()
        # ---- ragged gather, including an empty contribution
        t = torch.arange(3 * rank, dtype=torch.int64)            # rank 0: empty, rank 1: [0,1,2]
        cat, counts = sharded.gather_ragged(t)
        res["ragged"] = (cat.clone(), counts)
        # ---- scene merge, tiles sharded over ranks
        bx, sc, lb, ti, org = synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=120, seed=3)
        mb, ms, ml = sharded.scene_merge_nms(bx, sc, lb, ti, org, nms_fn=_oracle_nms)
        res["merge"] = (mb.clone(), ms.clone(), ml.clone())
        def to_np(x):        # plain numpy crosses the process boundary by value (torch tensors travel as fd handles)
            if isinstance(x, torch.Tensor):
                return x.numpy().copy()
            if isinstance(x, (tuple, list)):
                return tuple(to_np(y) for y in x)
            return x
        q.put((rank, {k: to_np(v) for k, v in res.items()}))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r, res = q.get(timeout=240)
        out[r] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    def to_t(x):
        if isinstance(x, np.ndarray):
            return torch.from_numpy(x)
        if isinstance(x, tuple):
            return tuple(to_t(y) for y in x)
        return x
    return {r: {k: to_t(v) for k, v in res.items()} for r, res in out.items()}


def test_row_sharded_iou_matches_single_process(two_ranks):
    from aidet_b200 import synth
    from oracle import oracle as O
    a, _ = synth.dota_boxes(37, side=200, seed=1)
    b, _ = synth.dota_boxes(53, side=200, seed=2)
    ref = torch.from_numpy(O.riou_matrix(a.numpy(), b.numpy()).astype(np.float32))
    for r in (0, 1):
        assert two_ranks[r]["iou"].shape == (37, 53)
        assert torch.equal(two_ranks[r]["iou"], ref)              # every rank holds the whole matrix
        part, r0, r1 = two_ranks[r]["iou_part"]
        assert (r0, r1) == ((0, 19) if r == 0 else (19, 37))
        assert torch.equal(part, ref[r0:r1])
        assert torch.equal(two_ranks[r]["iou_one"], ref[:1])      # 1 row over 2 ranks: rank 1 owns nothing


def test_ragged_gather(two_ranks):
    for r in (0, 1):
        cat, counts = two_ranks[r]["ragged"]
        assert list(counts) == [0, 3] and cat.tolist() == [0, 1, 2]


def test_scene_merge_is_world_size_independent(two_ranks):
    """2-rank result == single-process result == the straightforward two-stage oracle computation."""
    from aidet_b200 import sharded, synth
    from oracle import oracle as O
    bx, sc, lb, ti, org = synth.scene_dets(scene=1500, tile=512, overlap=100, dets_per_tile=120, seed=3)
    single = sharded.scene_merge_nms(bx, sc, lb, ti, org, nms_fn=_oracle_nms)          # no process group
    for r in (0, 1):
        for x, y in zip(two_ranks[r]["merge"], single):
            assert torch.equal(x, y)
    # independent restatement
    keep1, _ = O.nms(bx.numpy(), sc.numpy(), 0.5, groups=(ti * 15 + lb).int().numpy(), cmp_ge=False, plus_one=False)
    keep1 = torch.from_numpy(keep1)
    sb = sharded.translate_to_scene(bx[keep1], org[ti[keep1]])
    thr = sharded.merge_thresholds('obb').numpy()
    keep2, _ = O.nms(sb.numpy(), sc[keep1].numpy(), thr, groups=lb[keep1].int().numpy(), cmp_ge=False, plus_one=False)
    keep2 = torch.from_numpy(keep2)
    order = torch.argsort(lb[keep1][keep2], stable=True)
    assert torch.equal(single[0], sb[keep2][order])
    assert torch.equal(single[2], lb[keep1][keep2][order])
    assert len(keep2) < len(keep1) < len(bx)                     # both stages did suppress something
    # class thresholds follow mmdet/datasets/dota.py:324
    assert abs(float(sharded.merge_thresholds('obb')[sharded.DOTA_CLASSES.index('ship')]) - 0.05) < 1e-7
    assert float(sharded.merge_thresholds('hbb', classwise=False)[3]) == pytest.approx(0.3)


def test_shard_rows_and_translate():
    from aidet_b200 import sharded
    assert sharded.shard_rows(100000, 8, 7) == (12500, 87500, 100000)
    assert sharded.shard_rows(10, 4, 3) == (3, 9, 10)
    assert sharded.shard_rows(2, 4, 3) == (1, 2, 2)
    assert sharded.shard_rows(0, 4, 1) == (0, 0, 0)
    b5 = torch.tensor([[10.0, 20.0, 4.0, 2.0, 0.3]])
    b8 = torch.tensor([[0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]])
    o = torch.tensor([[100.0, 200.0]])
    assert sharded.translate_to_scene(b5, o).tolist() == [[110.0, 220.0, 4.0, 2.0, pytest.approx(0.3)]]
    assert sharded.translate_to_scene(b8, o).tolist() == [[100.0, 201.0, 102.0, 203.0, 104.0, 205.0, 106.0, 207.0]]
