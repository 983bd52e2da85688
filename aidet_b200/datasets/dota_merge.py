"""DOTA scene merge on the GPU -- drop-in for the merge step of `DOTADataset.evaluate`.

Reference call sites (mmdet/datasets/dota.py):
    :23       from wwtool.datasets.dota import mergebypoly, mergebypoly_mp, mergebyrec, mergebyrec_mp, ...
    :310-336  DOTADataset.merge_txt: per-class thresholds (:321-324), 0.3 everywhere when
              classwise_nms_threshold is False (:326-331), then
              mergebyrec_mp(txt_path, mergetxt_path, nms_thresh=hbb_nms_thr)   (task 'hbb', :334)
              mergebypoly_mp(txt_path, mergetxt_path, o_thresh=obb_nms_thr)    (task 'obb', :336)
    :278-308  format_dota_results: one `Task1_<class>.txt` / `Task2_<class>.txt` per class, rows
              `<tile> <score %.3f> x1 y1 ... x4 y4` (obb, %.1f) or `<tile> <score> x1 y1 x2 y2` (hbb)
    tools/dota/dota_demo.py:33  tile names `P0088__1.0__0___684` = <scene>__<rate>__<x>___<y>

`wwtool` is an un-vendored, un-pinned third-party package: its merge helpers are not in
/root/reference and cannot be run here, so the semantics below follow the public DOTA_devkit
`ResultMerge_multi_process.py` they derive from and are stated explicitly (parity for this file is UNPINNED):
    * a tile row is shifted by the tile origin and divided by the rate:  p_scene = (p_tile + (x, y)) / rate
    * detections are grouped per (class file, scene); greedy NMS per group in score order, a box is dropped when
      its overlap with an already kept box is  > thresh  (devkit: keeps `ovr <= thresh`)
    * obb: polygon IoU (the rotated-IoU kernel, 8-point form); hbb: axis-aligned IoU with the +1 convention
    * a merged file has one row per kept box `<scene> <score> <coords...>`, scenes in order of first
      appearance, boxes of a scene in score order (the devkit's keep order)
The reference runs this as a Python O(n^2) loop calling SWIG `polyiou` per pair inside a multiprocessing pool
over class files; here ALL class files and scenes go through ONE batched NMS launch
(groups = class x scene, per-group thresholds; aidet_nms_batched_f32).  `*_mp` names are kept for call
compatibility; there is no process pool.  With torch.distributed initialised, class files are dealt
round-robin to the ranks (each rank writes its own files; no data-path collective).

There is no CPU compute here: `nms_fn` exists so tests can exercise the host logic with the oracle.
"""
import os
import re
import shutil

import numpy as np
import torch

from ..sharded import DOTA_CLASSES, DOTA_HBB_MERGE_THR, DOTA_OBB_MERGE_THR, _world

_XY = re.compile(r'__(\d+)___(\d+)')
_RATE = re.compile(r'__([\d+\.]+)__\d+___')

TXT_SAVE_DIR = {'hbb': 'dota_hbb', 'obb': 'dota_obb'}                    # dota.py:61-62
MERGETXT_SAVE_DIR = {'hbb': 'merge_dota_hbb', 'obb': 'merge_dota_obb'}    # dota.py:63-64
TXT_FILE_PREFIX = {'hbb': 'Task2', 'obb': 'Task1'}                        # dota.py:65-66


def parse_tile_name(name):
    """`P0088__1.0__0___684` -> ('P0088', 1.0, 0, 684); a name without a tile suffix is its own scene at (0, 0)."""
    m = _XY.search(name)
    if m is None:
        return name, 1.0, 0, 0
    r = _RATE.search(name)
    rate = float(r.group(1)) if r else 1.0
    return name.split('__')[0], rate, int(m.group(1)), int(m.group(2))


def _read_task_file(path, ncoord):
    names, rows = [], []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) < 2 + ncoord:
                continue
            names.append(parts[0])
            rows.append([float(v) for v in parts[1:2 + ncoord]])
    arr = np.asarray(rows, dtype=np.float64).reshape(-1, 1 + ncoord)
    return names, arr[:, 0], arr[:, 1:]


def _nms_cuda(boxes, scores, groups, thr, n_groups, plus_one):
    from ..ops import functional as F
    return F.nms_batched(boxes, scores, groups, thr, n_groups=n_groups, cmp_ge=False, plus_one=plus_one)


def _merge(srcpath, dstpath, thresh, ncoord, plus_one, nms_fn=None, device=None, group=None):
    """Shared body of mergebypoly / mergebyrec.  Returns {class file name: number of rows written}."""
    world, rank = _world(group)
    files = sorted(f for f in os.listdir(srcpath) if f.endswith('.txt'))
    mine = [f for i, f in enumerate(files) if i % world == rank]
    os.makedirs(dstpath, exist_ok=True)
    names_all, scores_all, boxes_all, gid_all, thr_groups = [], [], [], [], []
    per_file = []                                   # (file, [scene names in order of first appearance], first group id)
    for fname in mine:
        names, scores, coords = _read_task_file(os.path.join(srcpath, fname), ncoord)
        cls = os.path.splitext(fname)[0].split('_', 1)[-1]          # Task1_<class>.txt
        if isinstance(thresh, dict):
            if cls not in thresh:
                raise KeyError("no merge threshold for class file %r" % fname)
            t = float(thresh[cls])
        else:
            t = float(thresh)
        scenes, scene_id = [], {}
        first = len(thr_groups)
        gids = np.empty((len(names),), np.int32)
        for i, nm in enumerate(names):
            scene, rate, x, y = parse_tile_name(nm)
            coords[i, 0::2] = (coords[i, 0::2] + x) / rate
            coords[i, 1::2] = (coords[i, 1::2] + y) / rate
            k = scene_id.get(scene)
            if k is None:
                k = scene_id[scene] = len(scenes)
                scenes.append(scene)
                thr_groups.append(t)
            gids[i] = first + k
        per_file.append((fname, scenes, first))
        scores_all.append(scores); boxes_all.append(coords); gid_all.append(gids)
    written = {}
    if not mine:
        return written
    scores_np = np.concatenate(scores_all) if scores_all else np.zeros((0,))
    boxes_np = np.concatenate(boxes_all) if boxes_all else np.zeros((0, ncoord))
    gids_np = np.concatenate(gid_all) if gid_all else np.zeros((0,), np.int32)
    n_groups = max(len(thr_groups), 1)
    if scores_np.shape[0]:
        fn = nms_fn or _nms_cuda
        dev = device if device is not None else (torch.device('cuda', torch.cuda.current_device()) if nms_fn is None
                                                 else torch.device('cpu'))
        keep = fn(torch.from_numpy(boxes_np.astype(np.float32)).to(dev), torch.from_numpy(scores_np.astype(np.float32)).to(dev),
                  torch.from_numpy(gids_np).to(dev), torch.tensor(thr_groups or [0.0], dtype=torch.float32, device=dev),
                  n_groups, plus_one)
        keep = keep.cpu().numpy()
    else:
        keep = np.zeros((0,), np.int64)
    # rows of a group in score order (stable: ties keep input order, like the sort inside the NMS)
    kg, ks = gids_np[keep], scores_np[keep]
    order = np.lexsort((keep, -ks.astype(np.float32), kg))
    keep, kg = keep[order], kg[order]
    bounds = np.searchsorted(kg, np.arange(n_groups + 1))
    for fname, scenes, first in per_file:
        rows = 0
        with open(os.path.join(dstpath, fname), 'w') as out:
            for k, scene in enumerate(scenes):
                for i in keep[bounds[first + k]:bounds[first + k + 1]]:
                    out.write(scene + ' ' + str(scores_np[i]) + ' ' + ' '.join(map(str, boxes_np[i])) + '\n')
                    rows += 1
        written[fname] = rows
    return written


def mergebypoly(srcpath, dstpath, o_thresh=0.3, nms_fn=None, device=None, group=None):
    """wwtool.datasets.dota.mergebypoly(srcpath, dstpath, o_thresh): polygon-NMS merge of `Task1_*.txt` (dota.py:336)."""
    return _merge(srcpath, dstpath, o_thresh, 8, False, nms_fn, device, group)


def mergebyrec(srcpath, dstpath, nms_thresh=0.3, nms_fn=None, device=None, group=None):
    """wwtool.datasets.dota.mergebyrec(srcpath, dstpath, nms_thresh): HBB (+1) NMS merge of `Task2_*.txt` (dota.py:334)."""
    return _merge(srcpath, dstpath, nms_thresh, 4, True, nms_fn, device, group)


mergebypoly_mp = mergebypoly      # the reference's *_mp variants fork a pool over class files; here one launch does all
mergebyrec_mp = mergebyrec


def merge_txt(submit_path, task='hbb', classwise_nms_threshold=True, nms_fn=None, device=None, group=None):
    """DOTADataset.merge_txt (dota.py:310-336) as a free function: same directories, same thresholds."""
    txt_path = os.path.join(submit_path, TXT_SAVE_DIR[task])
    mergetxt_path = os.path.join(submit_path, MERGETXT_SAVE_DIR[task])
    world, rank = _world(group)
    if rank == 0:
        if os.path.exists(mergetxt_path):
            shutil.rmtree(mergetxt_path)
        os.makedirs(mergetxt_path)
    if world > 1:
        torch.distributed.barrier(group)
    table = dict(DOTA_HBB_MERGE_THR if task == 'hbb' else DOTA_OBB_MERGE_THR)
    if not classwise_nms_threshold:
        table = {c: 0.3 for c in DOTA_CLASSES}
    if task == 'hbb':
        out = mergebyrec_mp(txt_path, mergetxt_path, nms_thresh=table, nms_fn=nms_fn, device=device, group=group)
    else:
        out = mergebypoly_mp(txt_path, mergetxt_path, o_thresh=table, nms_fn=nms_fn, device=device, group=group)
    if world > 1:
        torch.distributed.barrier(group)
    return out


def format_dota_results(submit_path, filenames, bboxes, scores, labels, task='hbb', classes=DOTA_CLASSES):
    """DOTADataset.format_dota_results (dota.py:278-308): rows of tile detections -> one txt per class.

    filenames[i] tile name, bboxes (n, 4) for 'hbb' / (n, 8) for 'obb', scores (n,), labels (n,) 1-based as in
    the reference (`CLASSES[labels[i] - 1]`).  (The reference's storage-tank special case, which swaps in the
    HBB corners, is dataset policy and is left to the caller.)
    """
    txt_path = os.path.join(submit_path, TXT_SAVE_DIR[task])
    if os.path.exists(txt_path):
        shutil.rmtree(txt_path)
    os.makedirs(txt_path)
    handles = {c: open(os.path.join(txt_path, "{}_{}.txt".format(TXT_FILE_PREFIX[task], c)), 'a+') for c in classes}
    try:
        for i, bbox in enumerate(bboxes):
            if task == 'hbb':
                row = '%s %.3f %.1f %.1f %.1f %.1f\n' % (filenames[i], scores[i], bbox[0], bbox[1], bbox[2], bbox[3])
            else:
                row = '%s %.3f %.1f %.1f %.1f %.1f %.1f %.1f %.1f %.1f\n' % (filenames[i], scores[i], bbox[0], bbox[1], bbox[2],
                                                                            bbox[3], bbox[4], bbox[5], bbox[6], bbox[7])
            handles[classes[int(labels[i]) - 1]].write(row)
    finally:
        for h in handles.values():
            h.close()
    return txt_path
