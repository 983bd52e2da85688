"""DOTA merge tool (aidet_b200.datasets.dota_merge) -- drop-in for DOTADataset.merge_txt
(mmdet/datasets/dota.py:310-336 -> wwtool mergebypoly_mp / mergebyrec_mp).

CPU tests exercise the host logic (txt grammar of dota.py:300-304, tile names of tools/dota/dota_demo.py:33,
scene grouping, per-class thresholds of dota.py:321-324, output order) with the float64 oracle plugged in as the
NMS callable and compare against an independent, devkit-style greedy loop written here.  The GPU test runs the
same files through the CUDA kernel."""
import os

import numpy as np
import pytest
import torch

from aidet_b200 import datasets as D
from aidet_b200 import sharded, synth
from oracle import oracle as O


def _oracle_nms(boxes, scores, groups, thr, n_groups, plus_one):
    keep, _ = O.nms(boxes.numpy(), scores.numpy(), thr.numpy(), groups=groups.numpy(), cmp_ge=False, plus_one=plus_one)
    return torch.from_numpy(keep)


def _write_scene(tmp_path, n_scenes=2, seed=11):
    """Two scenes of 3x3 tiles; returns submit_path and the rows written per task (after %.1f / %.3f rounding)."""
    names, obb, hbb, scores, labels = [], [], [], [], []
    for s in range(n_scenes):
        bx, sc, lb, ti, org = synth.scene_dets(scene=1300, tile=512, overlap=100, dets_per_tile=60, seed=seed + s)
        p8 = synth.thetaobb2pointobb(bx)
        for i in range(bx.size(0)):
            x0, y0 = org[ti[i]].tolist()
            names.append("P%04d__1.0__%d___%d" % (s, int(x0), int(y0)))
            q = p8[i].numpy()
            obb.append(q)
            hbb.append([q[0::2].min(), q[1::2].min(), q[0::2].max(), q[1::2].max()])
            scores.append(float(sc[i])); labels.append(int(lb[i]) + 1)
    submit = str(tmp_path)
    D.format_dota_results(submit, names, hbb, scores, labels, 'hbb')
    D.format_dota_results(submit, names, obb, scores, labels, 'obb')
    return submit


def _devkit_merge(path, ncoord, thr_table, plus_one):
    """Independent restatement of the devkit flow for ONE class file: parse, shift, group by scene, greedy loop."""
    cls = os.path.splitext(os.path.basename(path))[0].split('_', 1)[-1]
    thr = thr_table[cls]
    scenes = {}
    for line in open(path):
        p = line.split()
        scene, rate, x, y = D.parse_tile_name(p[0])
        c = np.array([float(v) for v in p[2:2 + ncoord]])
        c[0::2] = (c[0::2] + x) / rate
        c[1::2] = (c[1::2] + y) / rate
        scenes.setdefault(scene, []).append((float(p[1]), c))
    out = []
    for scene, dets in scenes.items():
        sc = np.array([d[0] for d in dets], np.float32)
        bx = np.array([d[1] for d in dets], np.float32)
        order = np.argsort(-sc, kind='stable')
        ovr = O.riou_matrix(bx, bx) if ncoord == 8 else O.hbb_overlaps(bx, bx, plus_one=plus_one)
        alive = np.ones(len(dets), bool)
        for a, i in enumerate(order):
            if not alive[i]:
                continue
            out.append((scene, i))
            for j in order[a + 1:]:
                if alive[j] and ovr[i, j] > thr:
                    alive[j] = False
    return out


@pytest.mark.parametrize("task", ["obb", "hbb"])
def test_merge_txt_host_logic(tmp_path, task):
    submit = _write_scene(tmp_path)
    written = D.merge_txt(submit, task, nms_fn=_oracle_nms)
    src = os.path.join(submit, D.dota_merge.TXT_SAVE_DIR[task])
    dst = os.path.join(submit, D.dota_merge.MERGETXT_SAVE_DIR[task])
    table = sharded.DOTA_OBB_MERGE_THR if task == 'obb' else sharded.DOTA_HBB_MERGE_THR
    ncoord = 8 if task == 'obb' else 4
    assert sorted(os.listdir(dst)) == sorted(os.listdir(src)) and len(os.listdir(dst)) == 15
    total_in = total_out = 0
    for f in os.listdir(src):
        n_in = sum(1 for _ in open(os.path.join(src, f)))
        rows = [l.split() for l in open(os.path.join(dst, f))]
        ref = _devkit_merge(os.path.join(src, f), ncoord, table, True)
        assert written[f] == len(rows) == len(ref) <= n_in
        assert [r[0] for r in rows] == [s for s, _ in ref]                     # scene order, score order inside
        assert all(len(r) == 2 + ncoord for r in rows)
        total_in += n_in; total_out += len(rows)
    assert 0 < total_out < total_in                                             # duplicates across tiles were merged


def test_merge_flat_threshold_and_names(tmp_path):
    submit = _write_scene(tmp_path, n_scenes=1)
    a = D.merge_txt(submit, 'obb', classwise_nms_threshold=False, nms_fn=_oracle_nms)
    src = os.path.join(submit, 'dota_obb')
    b = D.mergebypoly_mp(src, os.path.join(submit, 'again'), o_thresh={c: 0.3 for c in D.DOTA_CLASSES}, nms_fn=_oracle_nms)
    assert a == b
    c = D.mergebypoly(src, os.path.join(submit, 'scalar'), o_thresh=0.3, nms_fn=_oracle_nms)
    assert a == c
    with pytest.raises(KeyError):
        D.mergebypoly(src, os.path.join(submit, 'bad'), o_thresh={'plane': 0.3}, nms_fn=_oracle_nms)
    assert D.parse_tile_name('P0088__1.0__0___684') == ('P0088', 1.0, 0, 684)
    assert D.parse_tile_name('P1__0.5__824___1648') == ('P1', 0.5, 824, 1648)


def test_merge_rate_and_empty_file(tmp_path):
    src = tmp_path / 'src'; src.mkdir()
    # two sightings of one object from tiles at rate 0.5: tile coords are scene coords * rate - origin
    (src / 'Task1_plane.txt').write_text(
        "S__0.5__0___0 0.900 10.0 10.0 30.0 10.0 30.0 20.0 10.0 20.0\n"
        "S__0.5__8___0 0.800 2.0 10.0 22.0 10.0 22.0 20.0 2.0 20.0\n"
        "S__0.5__100___100 0.700 10.0 10.0 30.0 10.0 30.0 20.0 10.0 20.0\n")
    (src / 'Task1_ship.txt').write_text("")
    out = D.mergebypoly(str(src), str(tmp_path / 'dst'), o_thresh=0.3, nms_fn=_oracle_nms)
    assert out == {'Task1_plane.txt': 2, 'Task1_ship.txt': 0}
    rows = [l.split() for l in open(tmp_path / 'dst' / 'Task1_plane.txt')]
    assert rows[0][0] == 'S' and float(rows[0][1]) == pytest.approx(0.9)
    assert [float(v) for v in rows[0][2:]] == [20.0, 20.0, 60.0, 20.0, 60.0, 40.0, 20.0, 40.0]
    assert [float(v) for v in rows[1][2:4]] == [220.0, 220.0]
    assert open(tmp_path / 'dst' / 'Task1_ship.txt').read() == ""


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["obb", "hbb"])
def test_merge_txt_gpu_matches_oracle(cuda, tmp_path, task):
    submit = _write_scene(tmp_path, n_scenes=3, seed=21)
    got = D.merge_txt(submit, task)                                   # CUDA kernel
    dst = os.path.join(submit, D.dota_merge.MERGETXT_SAVE_DIR[task])
    gpu_rows = {f: open(os.path.join(dst, f)).read() for f in os.listdir(dst)}
    ref = D.merge_txt(submit, task, nms_fn=_oracle_nms)               # oracle as the checker
    ref_rows = {f: open(os.path.join(dst, f)).read() for f in os.listdir(dst)}
    assert got == ref
    assert gpu_rows == ref_rows
