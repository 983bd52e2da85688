#!/usr/bin/env python
"""bench.py -- throughput of the OBB hot path on B200 (one JSON line on stdout, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|iou|nms|roi]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): rotated IoU Gpairs/s, rotated-NMS Mboxes/s, RoIAlign GB/s.  The line's
top-level `metric`/`value` is the rotated-IoU matrix (first named, config C4: 100k x 100k, dense
synthetic set -- every pair truly intersects, so the bounding-circle early-out never fires and the
256 flop/pair charge is earned); the NMS (config C2) and RoIAlign (config C3) figures ride in the
`nms` and `roialign` objects of the same line, each with its own roofline, e2e and cpu_baseline.

A "step" is one pass of the op over one batch of synthetic input that is already resident in HBM;
`e2e` repeats it through the public Python API with pinned HOST buffers, H2D and D2H copies inside
the timed region.  N > 1: IoU rows are sharded over the ranks (strong scaling of the same 100k x
100k problem) and the shard results are exchanged with one in-place NCCL all-gather, which IS inside
the timed region (`compute_only` gives the figure without it); NMS groups and RoIs are sharded with
no data-path collective.

--impl reference times the CPU implementation of the same path on the host cores (the reference
has no in-tree rotated IoU/NMS/RoIAlign -- see DESIGN.md -- so this is the float64 oracle port under
oracle/, with OpenMP over all cores, on a bounded sample of the same workload; the reference's own
axis-aligned nms_cpu.cpp, compiled unmodified into oracle/_ref, is timed beside it).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_PAIR = 256.0                       # FP32 flop charged per box pair
FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4 TFLOP/s at the 1965 MHz maximum clock
SCALES = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
TRAFFIC = {}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="all", choices=["all", "iou", "nms", "roi"])
    p.add_argument("--iou-n", type=int, default=100000, help="IoU matrix side (config C4: 100000)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--quick", action="store_true", help="profiler runs: 5 graph replays instead of 200, no per-(image, level) "
                                                        "comparison loop in the RPN leg (thousands of launches under ncu)")
    return p.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the dominant kernels from the committed ncu capture (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed regions run (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # "under load" = samples in the upper half of the power range seen
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            load = [s for s, w in zip(sm, pw) if w >= thr] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), power_w_max=max(pw),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------ distributed plumbing
class Dist:
    def __init__(self, args):
        import torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.on = self.world > 1
        if self.on and args.impl == "ours":
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.on:
            self.dist.barrier()

    def max_float(self, x, device):
        import torch
        if not self.on:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_float(self, x, device):
        import torch
        if not self.on:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def finish(self):
        if self.on and hasattr(self, "dist"):
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed(D, dev, steps, warmup, fn, flush=None):
    """W untimed + K timed calls of fn(); barrier + synchronize on both sides; device time via CUDA
    events on the launching stream; returns (max-over-ranks total ms of the K steps, launches of ours)."""
    import torch
    from aidet_b200 import _lib as L
    for _ in range(warmup):
        if flush is not None:
            flush()
        fn()
    torch.cuda.synchronize(dev)
    D.barrier()
    torch.cuda.synchronize(dev)
    import gc
    gc.collect()
    gc_was = gc.isenabled()
    gc.disable()                                             # a collection inside a 70 us step would be charged to it
    l0 = L.launch_count()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = L.launch_count() - l0
    if gc_was:
        gc.enable()
    D.barrier()
    torch.cuda.synchronize(dev)
    return D.max_float(ms, dev), launches


def wall_timed(D, dev, steps, warmup, fn):
    """End-to-end legs: host wall clock around K calls that each end with the result on the host."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(dev)
    ms = (time.perf_counter() - t0) * 1e3
    D.barrier()
    return D.max_float(ms, dev)


# ------------------------------------------------------------------ workloads
def iou_config(n, G):
    """`config` of the headline line; the reference arm prints the identical dict (same workload, same keys)."""
    rows_per = (n + G - 1) // G
    return {"workload": "C4 rotated IoU matrix %dx%d theta-OBB (cx,cy,w,h,theta), dense synthetic set, rows sharded over "
                        "the GPUs + all-gather of the shard results when there are several" % (n, n),
            "iou_n": n, "box_format": "thetaobb (n,5)",
            "l2": "result %.1f GB per step >> 126 MB L2 (no flush needed); small legs flush L2 between steps" % (n * float(n) * 4 / 1e9)}


def iou_inputs(n, dense=True):
    from aidet_b200 import synth
    a, _ = synth.dota_boxes(n, side=16384, seed=4, dense=dense)
    b, _ = synth.dota_boxes(n, side=16384, seed=5, dense=dense)
    return a, b


def nms_inputs(dense=False, images=1):
    """Config C2 candidates: 2000 proposals x 15 classes, score > 0.05 (rbbox_nms.py:30); `images`
    tiles are batched as images x classes groups."""
    import torch
    from aidet_b200 import synth
    bs, ss, gs = [], [], []
    for im in range(images):
        mb, msc = synth.multiclass_dets(2000, 15, seed=2 + im, dense=dense)
        n, C = msc.shape[0], 15
        boxes = mb.view(n, C + 1, 5)[:, 1:]
        valid = (msc[:, 1:] > 0.05).t()
        lab, rows = valid.nonzero(as_tuple=True)
        bs.append(boxes[rows, lab]); ss.append(msc[:, 1:][rows, lab]); gs.append(lab.int() + 15 * im)
    return torch.cat(bs).contiguous(), torch.cat(ss).contiguous(), torch.cat(gs).contiguous(), 15 * images


def group_pairs(groups, n_groups):
    import torch
    cnt = torch.bincount(groups.long(), minlength=n_groups).double()
    return float((cnt * (cnt - 1) / 2).sum().item())


def roi_touched_bytes(rois, lvl, batch, C, tile=1024, out=7, grid=2, device=None):
    """|U_l| summed over the levels: distinct feature pixels touched by any bilinear tap of any RoI (BASELINE.md
    section 3), counted with torch ops on `device` -- the sampling rule of roi_align_kernel_v2.cu:79-126 (aligned
    variant, rotated about the RoI centre) restated here so that the product arm never touches oracle/."""
    import torch
    total = 0
    rois = rois.to(device).double()
    lvl = lvl.to(device)
    for l, s in enumerate(SCALES):
        r = rois[lvl == l]
        if r.numel() == 0:
            continue
        hw = int(tile * s)
        b = r[:, 0].long()
        cx, cy, w, h, th = r[:, 1] * s - 0.5, r[:, 2] * s - 0.5, r[:, 3] * s, r[:, 4] * s, r[:, 5]
        k = (torch.arange(out * grid, device=device, dtype=torch.float64) + 0.5) / (out * grid) - 0.5      # sample offsets / side
        yy = (h[:, None] * k[None, :])[:, :, None]
        xx = (w[:, None] * k[None, :])[:, None, :]
        cs, sn = torch.cos(th)[:, None, None], torch.sin(th)[:, None, None]
        x = cx[:, None, None] + xx * cs - yy * sn
        y = cy[:, None, None] + xx * sn + yy * cs
        ok = ~((y < -1.0) | (y > hw) | (x < -1.0) | (x > hw))
        x, y = x.clamp(min=0), y.clamp(min=0)
        xl, yl = x.floor().long().clamp(max=hw - 1), y.floor().long().clamp(max=hw - 1)
        xh, yh = (xl + 1).clamp(max=hw - 1), (yl + 1).clamp(max=hw - 1)
        touched = torch.zeros((batch * hw * hw,), dtype=torch.bool, device=device)
        base = (b * hw * hw)[:, None, None]
        for yy_, xx_ in ((yl, xl), (yl, xh), (yh, xl), (yh, xh)):
            idx = (base + yy_ * hw + xx_)[ok]
            touched[idx] = True
        total += int(touched.sum().item())
    return total


# ------------------------------------------------------------------ our arm
def run_ours(args, D):
    import torch
    from aidet_b200 import _lib as L, synth
    from aidet_b200.ops import functional as Fn
    dev = torch.device("cuda", D.local)
    torch.cuda.set_device(dev)
    hbm_peak, hbm_src = measured_peaks()
    global TRAFFIC
    TRAFFIC = measured_traffic()
    G, r = D.world, D.rank
    line = {}
    sampler = ClockSampler(D.local) if r == 0 else None
    L.prof_enable(True)
    ffma = L.ffma_peak_tflops(D.local, 8192)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def flush():
        flush_buf.zero_()

    # ---------------- rotated IoU (C4), rows sharded over the ranks
    if args.workload in ("all", "iou"):
        n = args.iou_n
        a, b = iou_inputs(n, dense=True)
        rows_per = (n + G - 1) // G
        m_pad = rows_per * G
        a_pad = torch.cat([a, a[: m_pad - n]]) if m_pad > n else a
        a_dev, b_dev = a_pad.to(dev), b.to(dev)
        my = a_dev[r * rows_per:(r + 1) * rows_per].contiguous()
        out = torch.empty((m_pad, n), dtype=torch.float32, device=dev)      # N=1: 40 GB >> L2, no flush needed
        shard = out[r * rows_per:(r + 1) * rows_per]

        from aidet_b200 import sharded

        def step_compute():
            Fn.riou_matrix(my, b_dev, out=shard)

        def step_full():         # public multi-GPU entry point: row shard + one in-place NCCL all-gather
            sharded.sharded_rbbox_overlaps(a_dev[:n], b_dev, out=out, gather=True)

        L.prof_read(L.PROF_RIOU, reset=True)
        ms, launches = timed(D, dev, args.steps, args.warmup, step_full)
        k_ms, k_cnt = L.prof_read(L.PROF_RIOU, reset=True)
        k_ms_avg = D.max_float(k_ms / max(k_cnt, 1), dev)
        pairs = float(n) * n
        ms_step = ms / args.steps
        comp_ms = ms_step
        multi = None
        if D.on:
            cms, _ = timed(D, dev, args.steps, args.warmup, step_compute)
            comp_ms = cms / args.steps
            multi = {"nccl_allgather": {"value": pairs / ms_step / 1e6, "unit": "Gpairs/s", "ms_per_step": ms_step,
                                        "how": "kernel on the row shard, then ONE in-place ncclAllGather of the shard results"}}
            # fused form: the kernel stores every tile to all ranks' symmetric-memory buffers over NVLink
            try:
                sym = sharded.SymmetricMatrix(n, n, dev)

                def step_fused():
                    sharded.sharded_rbbox_overlaps_fused(a_dev[:n], b_dev, sym)

                L.prof_read(L.PROF_RIOU, reset=True)
                fms, flaunches = timed(D, dev, args.steps, args.warmup, step_fused)
                fk_ms, fk_cnt = L.prof_read(L.PROF_RIOU, reset=True)
                fms_step = fms / args.steps
                # every rank must now hold the same matrix as the NCCL path produced
                step_full()
                step_fused()
                torch.cuda.synchronize(dev)
                probe = torch.arange(0, n, max(1, n // 64), device=dev)
                same = bool(torch.equal(sym.tensor[:n][probe], out[:n][probe]))
                same = D.max_float(0.0 if same else 1.0, dev) == 0.0
                # ... and the float64 oracle on 64 rows per rank, drawn from ALL shards (checker only: 64 x n pairs on the host)
                import numpy as np
                from oracle import oracle as O
                rs_ = np.sort(np.random.default_rng(100 + r).choice(n, 64, replace=False))
                got_ = sym.tensor[:n][torch.from_numpy(rs_).to(dev)].cpu().numpy().astype(np.float64)
                err_ = float(np.abs(got_ - O.riou_matrix(a.numpy()[rs_], b.numpy())).max())
                err_ = D.max_float(err_, dev)
                bytes_out = float(rows_per) * n * 4 * (G - 1)            # leaves this GPU over NVLink per step
                multi["fused_peer_stores"] = {
                    "value": pairs / fms_step / 1e6, "unit": "Gpairs/s", "ms_per_step": fms_step,
                    "kernel_ms": D.max_float(fk_ms / max(fk_cnt, 1), dev), "matches_nccl_path": same,
                    "max_abs_err": err_, "max_abs_err_how": "64 rows per rank (all shards) of the gathered matrix vs the float64 oracle; bound 1e-5",
                    "nvlink_bytes_out_per_rank": bytes_out, "nvlink_out_gbs": bytes_out / (fms_step * 1e-3) / 1e9,
                    "link_roofline": {"bound": "nvlink", "peak": 770.0, "unit": "GB/s per direction per GPU (measured peer copy, "
                                      "B200_PROFILING.md)", "frac": bytes_out / (fms_step * 1e-3) / 1e9 / 770.0,
                                      "target_ms": max(comp_ms, bytes_out / 770e9 * 1e3)},
                    "how": "aidet_riou_matrix_multi_f32: each tile is stored to the same rows of every rank's symmetric-memory "
                           "buffer from inside the kernel (NVLink peer stores), two symmetric-memory barriers, no NCCL on the data path"}
                if same and err_ <= 1e-5 and fms_step < ms_step:
                    ms_step, launches, k_ms_avg = fms_step, flaunches, D.max_float(fk_ms / max(fk_cnt, 1), dev)
                # NVSwitch multicast form: one multimem.st per element, replicated by the switch (NVLS)
                has_mc = D.max_float(0.0 if sym.mc_ptr else 1.0, dev) == 0.0
                if has_mc and os.environ.get("AIDET_BENCH_NO_MCAST") != "1":
                    def step_mcast():
                        sharded.sharded_rbbox_overlaps_fused(a_dev[:n], b_dev, sym, multicast=True)

                    sym.tensor.zero_()
                    L.prof_read(L.PROF_RIOU, reset=True)
                    mms, mlaunches = timed(D, dev, args.steps, args.warmup, step_mcast)
                    mk_ms, mk_cnt = L.prof_read(L.PROF_RIOU, reset=True)
                    mms_step = mms / args.steps
                    sym.tensor.zero_()
                    step_mcast()
                    torch.cuda.synchronize(dev)
                    same_mc = bool(torch.equal(sym.tensor[:n][probe], out[:n][probe]))
                    same_mc = D.max_float(0.0 if same_mc else 1.0, dev) == 0.0
                    own = float(rows_per) * n * 4
                    multi["fused_multicast_stores"] = {
                        "value": pairs / mms_step / 1e6, "unit": "Gpairs/s", "ms_per_step": mms_step,
                        "kernel_ms": D.max_float(mk_ms / max(mk_cnt, 1), dev), "matches_nccl_path": same_mc,
                        "nvlink_bytes_out_per_rank": own, "nvlink_bytes_in_per_rank": own * (G - 1),
                        "nvlink_in_gbs": own * (G - 1) / (mms_step * 1e-3) / 1e9,
                        "how": "aidet_riou_matrix_mcast_f32: every element stored once with multimem.st to the symmetric "
                               "buffer's multicast address; the NVSwitch replicates it into all ranks (this one included)"}
                    if same_mc and mms_step < ms_step:
                        ms_step, launches, k_ms_avg = mms_step, mlaunches, D.max_float(mk_ms / max(mk_cnt, 1), dev)
                else:
                    multi["fused_multicast_stores"] = {"unavailable": "no multicast mapping for the symmetric allocation"}
                del sym
            except Exception as exc:       # symmetric memory unavailable on this box: the NCCL number stands
                multi["fused_peer_stores"] = {"unavailable": repr(exc)[:300]}
        kernel_pairs = float(rows_per) * n                                   # per launch, per rank
        achieved = kernel_pairs * F_PAIR / (k_ms_avg * 1e-3) / 1e12
        line.update({
            "metric": "rotated IoU Gpairs/s", "value": pairs / ms_step / 1e6, "unit": "Gpairs/s",
            "n_gpus": G, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # compute_only / link bound repeated inside `config` (a key the driver's parser keeps): the gathered matrix is
            # bound by what every GPU must RECEIVE over NVLink ((G-1)/G of 4 n^2 bytes at the measured 770 GB/s per direction),
            # whatever the kernel does; the arithmetic alone scales linearly
            "config": dict(iou_config(n, G), rows_per_rank=rows_per, n_gpus=G,
                           compute_only_gpairs_s=pairs / comp_ms / 1e6,
                           **({"gather_link_bound_gpairs_s": pairs / (float(rows_per) * n * 4 * (G - 1) / 770e9) / 1e9,
                               "gather_link_bound_how": "pairs / ((G-1)/G x 4 n^2 bytes received per GPU / 770 GB/s)"} if G > 1 else {})),
            "compute_only": {"value": pairs / comp_ms / 1e6, "unit": "Gpairs/s", "ms_per_step": comp_ms},
            "multi_gpu": multi,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32-alu", "achieved": achieved, "peak": FP32_PEAK_NOMINAL, "unit": "TFLOP/s",
                         "frac": achieved / FP32_PEAK_NOMINAL,
                         "traffic": (TRAFFIC["riou_matrix_kernel"]["bytes_per_pair"] * kernel_pairs
                                     if "riou_matrix_kernel" in TRAFFIC else None),
                         "traffic_source": "ncu dram bytes per pair (%s) x pairs per launch" % TRAFFIC.get("source"),
                         "kernel": "riou_matrix_kernel<RectKind>", "kernel_ms": k_ms_avg,
                         "flop_per_pair": F_PAIR, "pairs_per_launch": kernel_pairs,
                         "peak_source": "148 SM x 128 lanes x 2 x 1.965 GHz (nominal max clock); FFMA micro-benchmark "
                                        "on this box: %.1f TFLOP/s" % ffma,
                         "frac_of_measured_ffma": achieved / ffma,
                         "hbm_write_gbs": kernel_pairs * 4 / (k_ms_avg * 1e-3) / 1e9},
        })
        # DOTA-shaped (sparse) set: pairs/s only, no roofline fraction
        ns = min(n, 32768)
        sa, sb = iou_inputs(ns, dense=False)
        sa, sb = sa.to(dev), sb.to(dev)
        rs = (ns + G - 1) // G
        so = out.view(-1)[: rs * ns].view(rs, ns)
        sm_, _ = timed(D, dev, args.steps, args.warmup, lambda: Fn.riou_matrix(sa[r * rs:(r + 1) * rs], sb, out=so))
        line["dota_shaped"] = {"value": float(ns) * ns / (sm_ / args.steps) / 1e6, "unit": "Gpairs/s",
                               "workload": "%dx%d DOTA-shaped (clustered, sparse) boxes, compute only" % (ns, ns)}

        # 8-point (point-OBB) inputs of the same dense set, 32768 x 32768: rectangles given by their corners (what
        # thetaobb2pointobb / hobb2pointobb / the DOTA txt rows hold) and free convex quads (corners jittered by 3 % of
        # the box size: what a point-OBB head regresses), same 256 flop/pair charge
        q8 = {}
        da8, db8 = iou_inputs(ns, dense=True)
        sets8 = (("rect_as_8pt", synth.thetaobb2pointobb(da8).float(), synth.thetaobb2pointobb(db8).float()),
                 ("free_quads", synth.free_quads(da8, rel_noise=0.03, seed=7)[0].float(),
                  synth.free_quads(db8, rel_noise=0.03, seed=8)[0].float()))
        for tag8, qa, qb in sets8:
            qa, qb = qa.to(dev), qb.to(dev)
            L.prof_read(L.PROF_RIOU, reset=True)
            qm, _ = timed(D, dev, args.steps, args.warmup, lambda: Fn.riou_matrix(qa[r * rs:(r + 1) * rs], qb, out=so))
            qk, qc = L.prof_read(L.PROF_RIOU, reset=True)
            qk_avg = D.max_float(qk / max(qc, 1), dev)
            q_ach = float(rs) * ns * F_PAIR / (qk_avg * 1e-3) / 1e12
            q8[tag8] = {"value": float(ns) * ns / (qm / args.steps) / 1e6, "unit": "Gpairs/s", "kernel_ms": qk_avg,
                        "roofline_frac": q_ach / FP32_PEAK_NOMINAL}
        q8["workload"] = "%dx%d dense set as 8-point boxes, compute only, rows sharded over the GPUs" % (ns, ns)
        line["iou_8point"] = q8

        if not args.no_e2e:
            # e2e: boxes in pinned host memory -> H2D -> kernel -> D2H of the result, streamed in row
            # chunks through two pinned staging buffers (the full fp32 result is 40 GB).
            chunk = max(1, min(rows_per, (1 << 30) // (4 * n)))              # ~1 GiB of result per chunk
            a_host = my.cpu().pin_memory()
            b_host = b.pin_memory()
            stage = [torch.empty((chunk, n), dtype=torch.float32).pin_memory() for _ in range(2)]
            dbuf = [torch.empty((chunk, n), dtype=torch.float32, device=dev) for _ in range(2)]
            copy_stream = torch.cuda.Stream(dev)
            done = [torch.cuda.Event(), torch.cuda.Event()]
            checksum = [0.0]

            def e2e_step():
                ad = a_host.to(dev, non_blocking=True)
                bd = b_host.to(dev, non_blocking=True)
                cur = torch.cuda.current_stream(dev)
                for ci, r0 in enumerate(range(0, rows_per, chunk)):
                    k = ci & 1
                    rr = min(chunk, rows_per - r0)
                    done[k].synchronize()                                     # staging buffer k is free again
                    Fn.riou_matrix(ad[r0:r0 + rr], bd, out=dbuf[k][:rr])
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(ev)
                        stage[k][:rr].copy_(dbuf[k][:rr], non_blocking=True)
                        done[k].record(copy_stream)
                copy_stream.synchronize()
                checksum[0] = float(stage[0][0, 0])

            ems = wall_timed(D, dev, max(2, min(args.steps, 3)), 1, e2e_step)
            esteps = max(2, min(args.steps, 3))
            line["e2e"] = {"value": pairs / (ems / esteps) / 1e6, "unit": "Gpairs/s",
                           "h2d_bytes_per_step": int((rows_per + n) * 5 * 4),
                           "d2h_bytes_per_step": int(rows_per * n * 4), "ms_per_step": ems / esteps,
                           "steps": esteps,
                           "how": "pinned host boxes -> H2D -> riou_matrix per %d-row chunk -> D2H into pinned staging "
                                  "(double-buffered, copy stream); per rank: its row shard" % chunk}
            del stage, dbuf
        del out, shard
        torch.cuda.empty_cache()

    # ---------------- batched rotated NMS (C2)
    if args.workload in ("all", "nms"):
        nms = {}
        for tag, dense, images, fmt8 in (("c2", False, 1, False), ("c2_dense", True, 1, False), ("c2x8", False, 8, False), ("c2x8_dense", True, 8, False),
                                         ("c2_8point", False, 1, True), ("c2_dense_8point", True, 1, True)):
            cb, cs, cg, ng = nms_inputs(dense=dense, images=images)
            if fmt8:
                cb = synth.thetaobb2pointobb(cb).float()
            # shard groups over ranks (independent units, no data-path collective)
            mine = (cg % G) == r if G > 1 else torch.ones_like(cg, dtype=torch.bool)
            cbd, csd, cgd = cb[mine].to(dev), cs[mine].to(dev), cg[mine].to(dev)
            Fn.nms_batched(cbd, csd, cgd, 0.5, n_groups=ng)      # one-time costs (first cooperative launch of the process,
            torch.cuda.synchronize(dev)                          # workspace growth) stay out of the device-time average
            L.prof_read(L.PROF_NMS_MASK, reset=True)
            ms, launches = timed(D, dev, args.steps, args.warmup,
                                 lambda: Fn.nms_batched(cbd, csd, cgd, 0.5, n_groups=ng), flush=flush)
            k_ms, k_cnt = L.prof_read(L.PROF_NMS_MASK, reset=True)
            k_avg = D.max_float(k_ms / max(k_cnt, 1), dev)
            nb = D.sum_float(float(cbd.shape[0]), dev)
            gp = group_pairs(cg[mine], ng)
            ach = gp * F_PAIR / (k_avg * 1e-3) / 1e12
            nms[tag] = {"value": nb / (ms / args.steps) / 1e3, "unit": "Mboxes/s", "ms_per_step": ms / args.steps,
                        "boxes": int(nb), "groups": ng, "gpu_launches": int(launches),
                        "roofline": {"bound": "fp32-alu", "achieved": ach, "peak": FP32_PEAK_NOMINAL,
                                     "unit": "TFLOP/s", "frac": ach / FP32_PEAK_NOMINAL,
                                     "kernel": "whole call on the device (fused rank+mask+scan+compact kernel for n <= 8192; "
                                               "sort + mask + scan kernels above)",
                                     "kernel_ms": k_avg, "charged_pairs_per_launch": gp, "traffic": None}}
        # roofline config: ONE group of 16384 dense boxes (134 M pairs).  A single group does not shard (the greedy scan
        # is sequential in score order): at N > 1 every rank runs the same group -- replicas only, the figure is per replica
        for tag, fmt8 in (("one_group_dense", False), ("one_group_dense_8point", True)):
            bb, bs_ = synth.dota_boxes(16384, side=16384, seed=11, dense=True)
            if fmt8:
                bb = synth.thetaobb2pointobb(bb).float()
            bbd, bsd = bb.to(dev), bs_.to(dev)
            Fn.nms_batched(bbd, bsd, None, 0.5)
            torch.cuda.synchronize(dev)
            L.prof_read(L.PROF_NMS_MASK, reset=True)
            ms, launches = timed(D, dev, args.steps, args.warmup, lambda: Fn.nms_batched(bbd, bsd, None, 0.5), flush=flush)
            k_ms, k_cnt = L.prof_read(L.PROF_NMS_MASK, reset=True)
            k_avg = D.max_float(k_ms / max(k_cnt, 1), dev)
            gp = bbd.shape[0] * (bbd.shape[0] - 1) / 2.0
            ach = gp * F_PAIR / (k_avg * 1e-3) / 1e12
            nms[tag] = {"value": float(bbd.shape[0]) / (ms / args.steps) / 1e3, "unit": "Mboxes/s",
                        "ms_per_step": ms / args.steps, "boxes": int(bbd.shape[0]), "groups": 1,
                        "gpu_launches": int(launches), "scaling": "replicas only (per replica)",
                        "roofline": {"bound": "fp32-alu", "achieved": ach, "peak": FP32_PEAK_NOMINAL, "unit": "TFLOP/s",
                                     "frac": ach / FP32_PEAK_NOMINAL, "kernel": "whole call on the device (sort + mask + scan + compact)",
                                     "kernel_ms": k_avg, "charged_pairs_per_launch": gp, "traffic": None}}
        # config C5: 4000^2 scene, 25 tiles (1024, overlap 200), per-tile NMS + cross-tile class-wise merge; tiles and
        # merge classes sharded over the ranks, survivors exchanged with two small all-gathers
        from aidet_b200 import sharded
        sx, ssc, sl, st, so = synth.scene_dets()
        sxd, sscd, sld, std_, sod = sx.to(dev), ssc.to(dev), sl.to(dev), st.to(dev), so.to(dev)
        ms, launches = timed(D, dev, args.steps, args.warmup,
                             lambda: sharded.scene_merge_nms(sxd, sscd, sld, std_, sod), flush=flush)
        merged = sharded.scene_merge_nms(sxd, sscd, sld, std_, sod)
        nms["c5_scene"] = {"value": sx.shape[0] / (ms / args.steps) / 1e3, "unit": "Mboxes/s", "ms_per_step": ms / args.steps,
                           "boxes": int(sx.shape[0]), "tiles": int(so.shape[0]), "kept": int(merged[0].shape[0]),
                           "gpu_launches": int(launches), "scaling": "strong",
                           "workload": "C5: 4000x4000 scene, 25 tiles of 1024 (overlap 200), 2000 dets/tile, 15 classes: per-tile "
                                       "NMS @0.5 + cross-tile merge with the class thresholds of dota.py:324"}
        nms["c5_scene"]["how"] = ("tiles (stage 1) and merge classes (stage 2) dealt to the ranks, one fixed-size all-reduce of the keep "
                                  "masks per stage; one scene is ~1 ms of sort / scan latency, so sharding ONE scene does not beat "
                                  "one GPU -- c5_scenes_per_rank is the form that scales")
        if G > 1:
            # a DOTA test set is hundreds of scenes: one scene per rank at a time, no data-path collective (weak scaling)
            ms, launches = timed(D, dev, args.steps, args.warmup,
                                 lambda: sharded.scene_merge_nms(sxd, sscd, sld, std_, sod, group=sharded.SOLO), flush=flush)
            nms["c5_scenes_per_rank"] = {"value": G * sx.shape[0] / (ms / args.steps) / 1e3, "unit": "Mboxes/s",
                                         "ms_per_step": ms / args.steps, "boxes": int(G * sx.shape[0]), "scenes": G,
                                         "gpu_launches": int(launches), "scaling": "weak",
                                         "workload": "C5, one whole scene per rank per step (every rank merges its own scene)"}
        nms["workload"] = ("C2: 2000 proposals x 15 classes, score>0.05 candidates, thr 0.5, one launch over all "
                           "classes (c2 = DOTA-shaped, c2_dense = every pair intersects, c2x8 = 8 tiles batched); "
                           "*_8point = the same boxes as corner lists; L2 flushed between steps; ms_per_step = whole call incl. the "
                           "count read-back, roofline.kernel_ms = all kernels of the call on the device")
        if G == 1:
            # Soft-NMS (nms_cpu.cpp:70-201 is the reference's only implementation): all 15 classes in one launch
            sd, sg, sng = soft_nms_inputs()
            sdd, sgd = sd.to(dev), sg.to(dev)
            ms, launches = timed(D, dev, args.steps, args.warmup,
                                 lambda: Fn.soft_nms_batched(sdd, sgd, 0.3, 1, 0.5, 0.05, n_groups=sng), flush=flush)
            srows, scnt = Fn.soft_nms_batched(sdd, sgd, 0.3, 1, 0.5, 0.05, n_groups=sng)
            shost = sd.pin_memory()
            sghost = sg.pin_memory()

            def soft_e2e():
                rws, cnt = Fn.soft_nms_batched(shost.to(dev, non_blocking=True), sghost.to(dev, non_blocking=True), 0.3, 1,
                                               0.5, 0.05, n_groups=sng)
                return rws.cpu(), cnt.cpu()
            sems = wall_timed(D, dev, args.steps, args.warmup, soft_e2e)
            nms["soft_nms"] = {"value": sd.shape[0] / (ms / args.steps) / 1e3, "unit": "Mboxes/s", "ms_per_step": ms / args.steps,
                               "boxes": int(sd.shape[0]), "groups": int(sng), "kept": int(srows.shape[0]),
                               "gpu_launches": int(launches),
                               "e2e": {"value": sd.shape[0] / (sems / args.steps) / 1e3, "unit": "Mboxes/s",
                                       "ms_per_step": sems / args.steps, "h2d_bytes_per_step": int(sd.numel() * 4 + sg.numel() * 4),
                                       "d2h_bytes_per_step": int(srows.numel() * 4 + scnt.numel() * 8)},
                               "workload": "Soft-NMS (linear, iou_thr 0.3, min_score 0.05) of the C2 candidates' AABB envelopes, "
                                           "15 class groups in one launch (one CTA per class); sequential in the selected box, "
                                           "so latency bound"}
        if not args.no_e2e:
            import numpy as np
            from aidet_b200.ops import batched_rnms
            cb, cs, cg, ng = nms_inputs(dense=False, images=1)
            hb, hs, hg = cb.pin_memory(), cs.pin_memory(), cg.pin_memory()

            def nms_e2e():
                k = batched_rnms(hb.to(dev, non_blocking=True), hs.to(dev, non_blocking=True),
                                 hg.to(dev, non_blocking=True), 0.5, n_groups=ng)
                return k.cpu()
            ems = wall_timed(D, dev, args.steps, args.warmup, nms_e2e)
            kk = nms_e2e()
            nms["e2e"] = {"value": cb.shape[0] / (ems / args.steps) / 1e3, "unit": "Mboxes/s",
                          "h2d_bytes_per_step": int(cb.numel() * 4 + cs.numel() * 4 + cg.numel() * 4),
                          "d2h_bytes_per_step": int(kk.numel() * 8 + 4), "ms_per_step": ems / args.steps,
                          "how": "c2: pinned host boxes/scores/groups -> batched_rnms -> keep indices on the host"}
        line["nms"] = nms

    # ---------------- rotated RoIAlign fwd + bwd (C3)
    if args.workload in ("all", "roi"):
        feats_h = synth.fpn_features()
        rois_h, lvl_h = synth.rotated_rois()
        feats = [f.to(dev) for f in feats_h]
        rois, lvl = rois_h.to(dev), lvl_h.to(dev)
        K, C = rois.shape[0], feats[0].shape[3]
        out = Fn.rroi_align_forward(feats, rois, SCALES, (7, 7), 2, 2, lvl)
        go = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(dev)
        grads = [torch.empty_like(f) for f in feats]
        touched = roi_touched_bytes(rois_h, lvl_h, feats_h[0].shape[0], C, device=dev)
        feat_elems = sum(f.numel() for f in feats)
        bytes_fwd = 4.0 * (C * touched + K * 49 * C + 6 * K)
        bytes_bwd = 4.0 * (K * 49 * C + C * touched + 6 * K) + 4.0 * feat_elems

        def fwd():
            Fn.rroi_align_forward(feats, rois, SCALES, (7, 7), 2, 2, lvl, out=out)

        ws_holder = [None]

        def bwd():          # deterministic gather backward: writes every gradient pixel once (zero-fill fused)
            ws_holder[0] = Fn.rroi_align_backward_gather(go, grads, rois, SCALES, 2, 2, lvl, workspace=ws_holder[0])

        def bwd_scatter():  # red.global.add.v4 scatter into zero-filled maps (the reference's formulation)
            for g_ in grads:
                g_.zero_()
            Fn.rroi_align_backward(go, grads, rois, SCALES, 2, 2, lvl)

        L.prof_read(L.PROF_ROI_FWD, reset=True)
        fms, fl = timed(D, dev, args.steps, args.warmup, fwd, flush=flush)
        kf, kfc = L.prof_read(L.PROF_ROI_FWD, reset=True)
        L.prof_read(L.PROF_ROI_BWD, reset=True)
        bms, bl = timed(D, dev, args.steps, args.warmup, bwd, flush=flush)
        kb, kbc = L.prof_read(L.PROF_ROI_BWD, reset=True)
        sms, _ = timed(D, dev, args.steps, args.warmup, bwd_scatter, flush=flush)
        L.prof_read(L.PROF_ROI_BWD, reset=True)
        fms, bms = fms / args.steps, bms / args.steps
        tot_bytes = (bytes_fwd + bytes_bwd) * G
        roi = {"value": tot_bytes / ((fms + bms) * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": fms + bms,
               "fwd": {"ms": fms, "gbs": bytes_fwd * G / (fms * 1e-3) / 1e9, "bytes": bytes_fwd,
                       "kernel_ms": kf / max(kfc, 1)},
               "bwd": {"ms": bms, "gbs": bytes_bwd * G / (bms * 1e-3) / 1e9, "bytes": bytes_bwd,
                       "kernel_ms": kb / max(kbc, 1),
                       "how": "gather backward (tap list bucketed per pixel by counting + scan, one warp per 4 pixels): every "
                              "gradient pixel written once, zero-fill fused, no atomics on the maps",
                       "scatter_variant_ms": sms / args.steps},
               "gpu_launches": int(fl + bl), "scaling": "weak (replicas: every rank runs the full C3 batch)",
               "workload": "C3: %d rotated RoIs (512/img x 8), 7x7, C=%d, FPN P2-P5 of a 1024 tile, NHWC fp32, "
                           "sample_num 2, all levels in one launch; features %.1f MB > L2 and L2 flushed between steps"
                           % (K, C, feat_elems * 4 / 1e6),
               "touched_feature_pixels": touched,
               "roofline": {"bound": "hbm", "achieved": (bytes_fwd + bytes_bwd) / ((fms + bms) * 1e-3) / 1e9,
                            "peak": hbm_peak, "unit": "GB/s",
                            "frac": (bytes_fwd + bytes_bwd) / ((fms + bms) * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": hbm_src,
                            "traffic": ((TRAFFIC["rroi_align_fwd_taplist_kernel"]["bytes"] + TRAFFIC["rroi_gather_kernel_bwd"]["bytes"])
                                        if "rroi_gather_kernel_bwd" in TRAFFIC else None),
                            "traffic_source": "ncu dram read+write of the fwd kernel + the gather-backward kernel, one C3 launch each (%s)" % TRAFFIC.get("source"),
                            "fwd_frac": bytes_fwd / (fms * 1e-3) / 1e9 / hbm_peak,
                            "bwd_frac": bytes_bwd / (bms * 1e-3) / 1e9 / hbm_peak,
                            # the same with the bytes the kernels really move (ncu DRAM read + write of the committed capture):
                            # the gather backward never writes the zero-fill the algorithmic formula charges
                            "fwd_frac_dram": (TRAFFIC["rroi_align_fwd_taplist_kernel"]["bytes"] / (fms * 1e-3) / 1e9 / hbm_peak
                                              if "rroi_align_fwd_taplist_kernel" in TRAFFIC else None),
                            "bwd_frac_dram": (TRAFFIC["rroi_gather_kernel_bwd"]["bytes"] / (bms * 1e-3) / 1e9 / hbm_peak
                                              if "rroi_gather_kernel_bwd" in TRAFFIC else None)}}
        if not args.no_e2e:
            fh = [f.pin_memory() for f in feats_h]
            rh, lh = rois_h.pin_memory(), lvl_h.pin_memory()
            goh = go.cpu().pin_memory()
            out_h = torch.empty(out.shape, dtype=torch.float32).pin_memory()
            gh = [torch.empty(f.shape, dtype=torch.float32).pin_memory() for f in feats_h]

            def roi_e2e():
                fd = [f.to(dev, non_blocking=True) for f in fh]
                rd, ld = rh.to(dev, non_blocking=True), lh.to(dev, non_blocking=True)
                gd = goh.to(dev, non_blocking=True)
                o = Fn.rroi_align_forward(fd, rd, SCALES, (7, 7), 2, 2, ld)
                out_h.copy_(o, non_blocking=True)
                gr = [torch.empty_like(f) for f in fd]
                ws_holder[0] = Fn.rroi_align_backward_gather(gd, gr, rd, SCALES, 2, 2, ld, workspace=ws_holder[0])
                for hbuf, g_ in zip(gh, gr):
                    hbuf.copy_(g_, non_blocking=True)
                torch.cuda.synchronize(dev)
            es = max(2, min(args.steps, 5))
            ems = wall_timed(D, dev, es, 1, roi_e2e)
            h2d = feat_elems * 4 + rois_h.numel() * 4 + lvl_h.numel() * 4 + go.numel() * 4
            d2h = out.numel() * 4 + feat_elems * 4
            roi["e2e"] = {"value": tot_bytes / (ems / es * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(h2d),
                          "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / es, "steps": es,
                          "how": "pinned host features+rois+grad_out -> H2D -> fwd -> D2H out; gather bwd -> D2H grads"}
        line["roialign"] = roi


    # ---------------- consumers of the overlap in training: fused max-IoU assignment, rotated IoU loss
    if args.workload in ("all", "iou"):
        from aidet_b200.core import MaxIoUAssigner, rbbox_overlaps
        from aidet_b200.models import riou_loss
        n_gt = 128                         # anchors of a 1024 tile (P2-P6, 3 per cell) x a crowded DOTA tile's truths
        _, gt, _, lab = synth.assign_case(1024, n_gt, seed=21)
        bx = synth.anchor_grid()
        n_anchor = bx.shape[0]
        bx, gt, lab = bx.to(dev), gt.to(dev), lab.to(dev)
        assigner = MaxIoUAssigner(0.7, 0.3, 0.3, True)

        def assign_fused():
            assigner.assign(bx, gt, None, lab)

        def assign_unfused():               # the reference's structure on the device: matrix, then the reductions
            assigner.assign_wrt_overlaps(rbbox_overlaps(gt, bx), lab)

        ams, al = timed(D, dev, args.steps, args.warmup, assign_fused, flush=flush)
        ums, _ = timed(D, dev, args.steps, args.warmup, assign_unfused, flush=flush)
        pr, tg = synth.regression_pairs(65536, seed=22)
        pr, tg = pr.to(dev), tg.to(dev)

        def loss_step():
            p_ = pr.clone().requires_grad_(True)
            riou_loss(p_, tg, reduction='sum').backward()

        lms, ll = timed(D, dev, args.steps, args.warmup, loss_step)
        line["assign"] = {"value": n_anchor * float(n_gt) / (ams / args.steps * 1e-3) / 1e9, "unit": "Gpairs/s",
                          "ms_per_step": ams / args.steps, "gpu_launches": int(al),
                          "unfused_ms_per_step": ums / args.steps,
                          "workload": "MaxIoUAssigner(0.7, 0.3, 0.3) on %d theta-OBB grid anchors (1024 tile, P2-P6, 3 ratios) x %d DOTA-shaped truths, "
                                      "labels gathered; fused = two recompute passes, no (k,n) matrix; unfused = overlap "
                                      "matrix (134 MB) + assign_wrt_overlaps; every rank runs the full case" % (n_anchor, n_gt)}
        line["riou_loss"] = {"value": 65536 / (lms / args.steps * 1e-3) / 1e6, "unit": "Mpairs/s",
                             "ms_per_step": lms / args.steps, "gpu_launches": int(ll),
                             "workload": "riou_loss forward + backward on 65536 aligned theta-OBB pairs (aligned overlap "
                                         "kernel + analytic-gradient kernel + the -log/clamp/sum torch ops around them)"}


    # ---------------- latency of the small configs: C1 and C2, eager and under CUDA-graph replay
    if args.workload in ("all", "nms", "iou"):
        from aidet_b200.core import rbbox_overlaps as _rov

        def graph_us(fn, iters=(5 if args.quick else 200)):
            """fn() must not touch the host.  -> (eager us per call, graph-replay us per call), device time."""
            L.prof_enable(False)                          # no event records inside a capture
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            eager = e0.elapsed_time(e1) / iters * 1e3
            gr = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread of a multi-rank run may query events while this thread captures
            with torch.cuda.graph(gr, capture_error_mode="thread_local"):
                fn()
            gr.replay()
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(iters):
                gr.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            L.prof_enable(True)
            return eager, e0.elapsed_time(e1) / iters * 1e3

        a1, s1 = synth.dota_boxes(2000, side=1024, seed=0)
        b1, _ = synth.dota_boxes(2000, side=1024, seed=1)
        a1, s1, b1 = a1.to(dev), s1.to(dev), b1.to(dev)
        c1_out = torch.empty((2000, 2000), device=dev)

        def c1_step():
            Fn.riou_matrix(a1, b1, out=c1_out)
            Fn.nms_batched(a1, s1, None, 0.1, n_groups=1, sync=False)

        cb2, cs2, cg2, ng2 = nms_inputs(dense=False, images=1)
        cb2, cs2, cg2 = cb2.to(dev), cs2.to(dev), cg2.to(dev)

        def c2_step():
            Fn.nms_batched(cb2, cs2, cg2, 0.5, n_groups=ng2, sync=False)

        try:
            c1_e, c1_g = graph_us(c1_step)
            c2_e, c2_g = graph_us(c2_step)
        except Exception as exc:                      # a failed capture must not cost the bench line
            L.prof_enable(True)
            c1_e = c1_g = c2_e = c2_g = None
            line["latency_error"] = repr(exc)[:200]
        line["latency"] = {
            "c1": {"eager_us": c1_e, "graph_us": c1_g,
                   "workload": "C1: rotated IoU matrix 2000x2000 theta-OBB + single-class rotated NMS @0.1 of the 2000 boxes "
                               "(3 kernels: prologue + matrix + the fused NMS launch), no host read-back; back-to-back calls, L2 warm"},
            "c2": {"eager_us": c2_e, "graph_us": c2_g, "boxes": int(cb2.shape[0]),
                   "workload": "C2: batched rotated NMS of the 15 classes (one cooperative launch), keep count left on the device"},
            "how": "CUDA events around %d back-to-back calls; graph = one torch.cuda.CUDAGraph replayed as often" % (5 if args.quick else 200)}

    # the driver's record keeps the standard keys only: the other two quantities of the metric (NMS Mboxes/s, RoIAlign
    # GB/s) and the exchange-free IoU figure ride along inside `roofline`, which it keeps whole
    if "roofline" in line:
        ops = {}
        if "compute_only" in line:
            ops["iou_compute_only_gpairs_s"] = line["compute_only"]["value"]
        if "iou_8point" in line:
            ops["iou_8point_rect_gpairs_s"] = line["iou_8point"]["rect_as_8pt"]["value"]
            ops["iou_8point_rect_frac"] = line["iou_8point"]["rect_as_8pt"]["roofline_frac"]
            ops["iou_8point_free_quads_gpairs_s"] = line["iou_8point"]["free_quads"]["value"]
            ops["iou_8point_free_quads_frac"] = line["iou_8point"]["free_quads"]["roofline_frac"]
        if "nms" in line:
            for k in ("c2", "c2_dense", "c2x8", "c2x8_dense", "one_group_dense", "c2_8point", "one_group_dense_8point"):
                if k in line["nms"]:
                    ops["nms_%s_mboxes_s" % k] = line["nms"][k]["value"]
                    ops["nms_%s_whole_call_frac" % k] = line["nms"][k]["roofline"]["frac"]
        if "latency" in line:
            ops["c1_graph_us"] = line["latency"]["c1"]["graph_us"]
            ops["c2_graph_us"] = line["latency"]["c2"]["graph_us"]
        if "roialign" in line:
            ops["roialign_gbs"] = line["roialign"]["value"]
            ops["roialign_frac"] = line["roialign"]["roofline"]["frac"]
            ops["roialign_fwd_ms"] = line["roialign"]["fwd"]["ms"]
            ops["roialign_bwd_ms"] = line["roialign"]["bwd"]["ms"]
            ops["roialign_fwd_frac"] = line["roialign"]["roofline"]["fwd_frac"]
            ops["roialign_bwd_frac_dram"] = line["roialign"]["roofline"]["bwd_frac_dram"]
        line["roofline"]["other_ops"] = ops
    line["clocks"] = sampler.stop() if sampler else None
    line["ffma_peak_tflops_measured"] = ffma
    return line, dev


# ------------------------------------------------------------------ CPU legs (oracle = checker / baseline only)
def cpu_iou(sample, threads=None):
    from oracle import oracle as O
    a, b = iou_inputs(sample, dense=True)
    an, bn = a.numpy(), b.numpy()
    O.riou_matrix(an[:64], bn[:64])
    t0 = time.perf_counter()
    O.riou_matrix(an, bn)
    dt = time.perf_counter() - t0
    return float(sample) * sample / dt / 1e9, dt


def cpu_nms(dense=False):
    """float64 oracle greedy NMS, the 15 class groups spread over the host threads (the reference's
    mergebypoly_mp likewise maps classes over a process pool, dota.py:336)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    cb, cs, cg, ng = nms_inputs(dense=dense, images=1)
    cbn, csn, cgn = cb.numpy(), cs.numpy(), cg.numpy()
    parts = [np.nonzero(cgn == g)[0] for g in range(ng)]
    O.nms(cbn[:32], csn[:32], 0.5, plus_one=False)

    def one(idx):
        return O.nms(cbn[idx], csn[idx], 0.5, cmp_ge=False, plus_one=False)[0]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        list(ex.map(one, parts))
    dt = time.perf_counter() - t0
    return cbn.shape[0] / dt / 1e6, dt, cbn.shape[0]


def cpu_nms_reference_hbb():
    """The reference's own nms_cpu.cpp (oracle/_ref, compiled unmodified) on the AABB envelopes, per class."""
    import torch
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        return None
    from aidet_b200 import synth
    cb, cs, cg, ng = nms_inputs(dense=False, images=1)
    p = synth.thetaobb2pointobb(cb).view(-1, 4, 2)
    hbb = torch.cat([p.min(1)[0], p.max(1)[0]], 1)
    dets = torch.cat([hbb, cs[:, None]], 1)
    parts = [dets[cg == g].contiguous() for g in range(ng)]
    mod.nms(parts[0][:16], 0.5)
    t0 = time.perf_counter()
    for d in parts:
        if d.shape[0]:
            mod.nms(d, 0.5)
    dt = time.perf_counter() - t0
    return dets.shape[0] / dt / 1e6, dt


def soft_nms_inputs():
    """C2 candidates as axis-aligned envelopes [x1,y1,x2,y2,score] + class ids (Soft-NMS is an HBB op in the reference)."""
    import torch
    from aidet_b200 import synth
    cb, cs, cg, ng = nms_inputs(dense=False, images=1)
    p = synth.thetaobb2pointobb(cb).view(-1, 4, 2)
    hbb = torch.cat([p.min(1)[0], p.max(1)[0]], 1)
    return torch.cat([hbb, cs[:, None]], 1).contiguous(), cg, ng


def cpu_soft_nms_reference():
    """The reference's own soft_nms_cpu_kernel (nms_cpu.cpp:70-201, oracle/_ref, unmodified), class by class as
    its Python loops call it (linear, iou_thr 0.3, min_score 0.05)."""
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        return None
    dets, cg, ng = soft_nms_inputs()
    parts = [dets[cg == g].contiguous() for g in range(ng)]
    mod.soft_nms(parts[0][:16].clone(), 0.3, 1, 0.5, 0.05)
    t0 = time.perf_counter()
    kept = 0
    for d in parts:
        if d.shape[0]:
            kept += mod.soft_nms(d.clone(), 0.3, 1, 0.5, 0.05).shape[0]
    dt = time.perf_counter() - t0
    return dets.shape[0] / dt / 1e6, dt, kept


def cpu_roi(sample_rois=512):
    """float64 oracle rotated RoIAlign fwd + bwd (OpenMP) on the first `sample_rois` RoIs of C3, level by level."""
    import numpy as np
    import torch
    from aidet_b200 import synth
    from oracle import oracle as O
    feats = synth.fpn_features()
    rois, lvl = synth.rotated_rois()
    rois, lvl = rois[:sample_rois], lvl[:sample_rois]
    C = feats[0].shape[3]
    go = np.random.RandomState(0).randn(sample_rois, 7, 7, C).astype(np.float32)
    touched = roi_touched_bytes(rois, lvl, feats[0].shape[0], C, device="cpu")
    t0 = time.perf_counter()
    for l, s in enumerate(SCALES):
        sel = (lvl == l).nonzero().flatten()
        if sel.numel() == 0:
            continue
        O.roi_align_fwd(feats[l].numpy(), rois[sel].numpy(), s, (7, 7), 2, O.ROI_V2_ALIGNED)
        O.roi_align_bwd(go[sel.numpy()], tuple(feats[l].shape), rois[sel].numpy(), s, 2, O.ROI_V2_ALIGNED)
    dt = time.perf_counter() - t0
    K = sample_rois
    by = 4.0 * (C * touched + K * 49 * C + 6 * K) * 2 + 4.0 * sum(f.numel() for f in feats)
    return by / dt / 1e9, dt


def cpu_roi_torchvision():
    """torchvision.ops.roi_align on the CPU (the implementation the reference itself offers,
    mmdet/ops/roi_align/roi_align.py:138-141) forward + autograd backward on ALL 4096 C3 RoIs, level by level as
    SingleRoIExtractor loops (single_level.py:96-107).  torchvision has no rotated RoIAlign: the RoIs are the C3 RoIs
    with theta = 0 (same centres and sizes, hence the same amount of sampling and of touched pixels)."""
    import torch
    import torchvision
    from aidet_b200 import synth
    feats = [f.permute(0, 3, 1, 2) for f in synth.fpn_features()]          # NCHW views of the NHWC maps (channels_last)
    rois, lvl = synth.rotated_rois()
    hbb = torch.cat([rois[:, :1], rois[:, 1:3] - rois[:, 3:5] / 2, rois[:, 1:3] + rois[:, 3:5] / 2], 1)
    C = feats[0].shape[1]
    xs = [f.clone().requires_grad_(True) for f in feats]
    t0 = time.perf_counter()
    outs = []
    for l, sc in enumerate(SCALES):
        sel = lvl == l
        if sel.any():
            outs.append(torchvision.ops.roi_align(xs[l], hbb[sel], (7, 7), sc, 2, aligned=True))
    torch.cat(outs).sum().backward()
    dt = time.perf_counter() - t0
    touched = roi_touched_bytes(torch.cat([rois[:, :5], torch.zeros(rois.shape[0], 1)], 1), lvl, feats[0].shape[0], C, device="cpu")
    K = rois.shape[0]
    by = 4.0 * (C * touched + K * 49 * C + 6 * K) * 2 + 4.0 * sum(f.numel() for f in feats)
    return by / dt / 1e9, dt, torch.get_num_threads()


def cpu_baselines(args):
    cores = os.cpu_count()
    out = {}
    if args.workload in ("all", "iou"):
        v, dt = cpu_iou(8192)
        out["iou"] = {"value": v, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                      "sample": "8192x8192 block of the same dense set, float64 oracle (oracle/oracle_geom.c), OpenMP, %.1f s" % dt}
    if args.workload in ("all", "nms"):
        v, dt, nb = cpu_nms(False)
        out["nms"] = {"value": v, "unit": "Mboxes/s", "cores": min(cores, 15), "kind": "port",
                      "sample": "full C2 (%d candidates, 15 classes over a thread pool), float64 oracle, %.2f s" % (nb, dt)}
        ref = cpu_nms_reference_hbb()
        if ref is not None:
            out["nms_hbb_reference"] = {"value": ref[0], "unit": "Mboxes/s", "cores": 1, "kind": "reference",
                                        "sample": "reference nms_cpu.cpp (unmodified, oracle/_ref) on the AABB envelopes "
                                                  "of C2, class by class, %.3f s" % ref[1]}
        sref = cpu_soft_nms_reference()
        if sref is not None:
            out["soft_nms_reference"] = {"value": sref[0], "unit": "Mboxes/s", "cores": 1, "kind": "reference", "kept": sref[2],
                                         "sample": "reference soft_nms_cpu_kernel (nms_cpu.cpp:70-201, unmodified, oracle/_ref) on "
                                                   "the AABB envelopes of C2, class by class, linear, %.3f s" % sref[1]}
    if args.workload in ("all", "roi"):
        v, dt = cpu_roi(512)
        out["roialign"] = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                           "sample": "first 512 of the 4096 C3 RoIs, fwd+bwd, float64 oracle (oracle/oracle_roi.c), OpenMP, %.1f s" % dt}
        try:
            tv, tdt, tthreads = cpu_roi_torchvision()
            out["roialign_torchvision"] = {"value": tv, "unit": "GB/s", "cores": tthreads, "kind": "reference",
                                           "sample": "torchvision.ops.roi_align CPU forward + autograd backward, float32, all 4096 C3 "
                                                     "RoIs at theta = 0, per level (the CPU implementation the reference endorses, "
                                                     "roi_align.py:138-141), %.1f s" % tdt}
        except Exception as exc:                     # torchvision missing / too old on the box: the port figure stands
            out["roialign_torchvision"] = {"unavailable": repr(exc)[:200]}
    return out


def run_reference(args):
    """--impl reference: the CPU path on the host cores, same metric/config keys as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is rank 0 alone and may use every host thread
    # (the OpenMP runtime reads the variable when the oracle library is first loaded, which is below)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    n = args.iou_n
    sample = 4096
    for _ in range(args.warmup):
        cpu_iou(1024)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, _dt = cpu_iou(sample)
        vals.append(v)
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    value = statistics.median(vals)
    line = {"impl": "reference", "metric": "rotated IoU Gpairs/s", "value": value, "unit": "Gpairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(iou_config(n, args.gpus), rows_per_rank=(n + args.gpus - 1) // args.gpus, n_gpus=args.gpus),
            "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                             "sample": "each step = a %dx%d block of the config's %dx%d matrix (a rate, so the block size does not "
                                       "enter), float64 oracle port (the reference has no in-tree rotated IoU; wwtool/polyiou is "
                                       "un-vendored), OpenMP over all cores" % (sample, sample, n, n)},
            "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload in ("all", "nms"):
        v, dt, nb = cpu_nms(False)
        line["nms"] = {"value": v, "unit": "Mboxes/s", "cores": min(cores, 15), "kind": "port",
                       "sample": "full C2, %d candidates" % nb, "ms_per_step": dt * 1e3}
        ref = cpu_nms_reference_hbb()
        if ref is not None:
            line["nms"]["hbb_reference"] = {"value": ref[0], "unit": "Mboxes/s", "cores": 1, "kind": "reference"}
    if args.workload in ("all", "roi"):
        v, dt = cpu_roi(512)
        line["roialign"] = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                            "sample": "first 512 C3 RoIs fwd+bwd", "ms_per_step": dt * 1e3}
    return line


def main():
    # stdout carries exactly one JSON line: libraries that print there (NCCL version banner, ...) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    args = parse()
    if args.impl == "reference":
        line = run_reference(args)
        if line is not None:
            emit(line)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: aidet_b200 has no CPU path (use --impl reference for the CPU arm)")
    D = Dist(args)
    line, dev = run_ours(args, D)
    if D.rank == 0 and not args.no_cpu and D.world == 1:
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
        cb = cpu_baselines(args)
        if "iou" in cb:
            line["cpu_baseline"] = cb.pop("iou")
        for k, v in cb.items():
            tgt = {"nms": "nms", "nms_hbb_reference": "nms", "soft_nms_reference": "nms", "roialign": "roialign",
                   "roialign_torchvision": "roialign"}[k]
            if tgt in line:
                if k == "soft_nms_reference":
                    if "soft_nms" in line[tgt]:
                        line[tgt]["soft_nms"]["cpu_baseline"] = v
                    continue
                line[tgt][{"nms_hbb_reference": "cpu_baseline_hbb_reference",
                           "roialign_torchvision": "cpu_baseline_torchvision"}.get(k, "cpu_baseline")] = v
    if D.rank == 0:
        emit(line)
    D.finish()


if __name__ == "__main__":
    main()
