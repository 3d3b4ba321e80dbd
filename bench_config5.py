"""Part of bench.py (kept outside the product package: its CPU legs execute oracle/).  BASELINE.json config 5 through bench.py (`--workload config5`): 3-D NMS and RoI crop-resize on 10 000 random proposals
over a 256^3 map, one GPU (SURVEY.md 8d recipe: centres U(0,256)^3, sides U(16,128)^3, clipped, scores U(0,1),
default_rng(0)).  CUDA-event medians with the L2 flushed between iterations; the CPU restatement (`--impl reference`, and
the cpu_baseline leg of the GPU arm) is the oracle's numpy NMS (reference utils.py:122-157) and crop + trilinear resize
(reference model.py:265-289) on a bounded sample.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
NMS_SETTINGS = ((0.7, 500), (0.7, 10000), (0.3, 10000))


def config5_boxes(n=10000):
    rng = np.random.default_rng(0)
    c = rng.uniform(0, 256, size=(n, 3))
    s = rng.uniform(16, 128, size=(n, 3))
    b = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.float32)
    sc = rng.uniform(0, 1, size=n).astype(np.float32)
    return b, sc


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cfun_oracle as O
    return O


def cpu_rows(roi_sample=200):
    """the reference's own CPU path for the same inputs: numpy NMS in full (it finishes in about a second), crop-resize on the
    first `roi_sample` boxes scaled to 10 000"""
    import torch
    O = _oracle()
    b, sc = config5_boxes()
    rows = {}
    for thr, mx in NMS_SETTINGS:
        t0 = time.time()
        keep = O.non_max_suppression(b, sc, thr, mx)
        rows["nms_thr%.1f_max%d" % (thr, mx)] = {"ms": (time.time() - t0) * 1e3, "kept": int(len(keep))}
    g = torch.Generator().manual_seed(11)
    fmap = torch.randn(1, 256, 256, 256, generator=g)
    boxes = torch.from_numpy(b / 256.0)
    t0 = time.time()
    O.roi_align(fmap, (12, 12, 12), boxes[:roi_sample])
    rows["roi_c1_pool12_10k"] = {"ms": (time.time() - t0) * 1e3 * (10000.0 / roi_sample),
                                 "sample": "%d of 10000 boxes, scaled" % roi_sample}
    return rows


def run_config5(args, peaks):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    name = ("RoIAlign3D + 3D-NMS microbench: 10k random 3D proposals over a 256^3 feature map, 1 GPU "
            "(NMS thr/max = 0.7/500, 0.7/all, 0.3/all; crop-resize C=1 256^3 pool 12^3 x 10000, pool 96^3 x 16, C=128 32^3 pool 12^3 x 1000)")
    if args.impl == "reference":
        import torch
        torch.set_num_threads(cores)
        rows = cpu_rows()
        total = sum(r["ms"] for r in rows.values())
        line = {"impl": "reference", "metric": "config5_microbench_passes_per_sec", "value": 1000.0 / total, "unit": "passes/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": total, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": name}, "rows": rows,
                "cpu_baseline": {"value": 1000.0 / total, "unit": "passes/s", "cores": cores, "kind": "port",
                                 "sample": "3 NMS settings in full + crop-resize of 200 of the 10000 boxes scaled up"},
                "e2e": {"value": 1000.0 / total, "unit": "passes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    import torch
    from cfun_b200 import ops, utils as U
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def med(fn, iters=max(5, args.steps)):
        for _ in range(max(3, args.warmup)):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        return sorted(ts)[len(ts) // 2]

    b, sc = config5_boxes()
    n = b.shape[0]
    bd, sd = torch.from_numpy(b).to(dev), torch.from_numpy(sc).to(dev)
    rows = {}
    l0 = ops.launch_count()
    for thr, mx in NMS_SETTINGS:
        def run():
            order = ops.sort_desc(sd)
            return ops.nms3d(bd[order.long()], thr, mx)
        t = med(run)
        keep, cnt = run()
        kept = int(cnt)
        def run_e2e():       # the public call: host numpy boxes in, kept indices out (H2D + D2H inside)
            return U.non_max_suppression(b, sc, thr, mx)
        t0 = time.time()
        for _ in range(3):
            got = run_e2e()
        te = (time.time() - t0) / 3 * 1e3
        alg = n * 7 * 4 + 4 * kept          # read boxes + scores once, write the kept indices
        rows["nms_thr%.1f_max%d" % (thr, mx)] = {
            "ms": t, "kept": kept, "e2e_ms_host_to_host": te, "iou_evaluations_per_s": (float(n) * max(kept, 1)) / (t * 1e-3),
            "algorithmic_bytes": alg, "gbs": alg / (t * 1e-3) / 1e9, "bound": "latency (O(n x kept) IoUs, sequential scan over 64-box blocks)"}
    boxes = torch.from_numpy(b / 256.0).to(dev)
    fmap = torch.randn(1, 1, 256, 256, 256, device=dev)
    f2 = torch.randn(1, 128, 32, 32, 32, device=dev)
    for key, fm, bx, pool, C in (("roi_c1_pool12_10k", fmap, boxes, 12, 1), ("roi_c1_pool96_16", fmap, boxes[:16], 96, 1),
                                 ("roi_c128_pool12_1k", f2, boxes[:1000], 12, 128)):
        t = med(lambda: ops.roi_crop_resize(fm, None, bx, None, (pool,) * 3, True))
        nb = bx.shape[0]
        out_b = nb * C * pool ** 3 * 4.0
        in_b = nb * C * 8.0 * pool ** 3 * 4.0           # gather-minimal: the 8 source voxels of every output voxel
        rows[key] = {"ms": t, "boxes": nb, "algorithmic_bytes": in_b + out_b, "gbs": (in_b + out_b) / (t * 1e-3) / 1e9,
                     "frac_of_hbm_peak": (in_b + out_b) / (t * 1e-3) / 1e9 / peaks["hbm_gbs"], "bound": "hbm (gather)"}
    launches = ops.launch_count() - l0
    total = sum(r["ms"] for r in rows.values())
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(cores)
        cr = cpu_rows()
        cpu = {"value": 1000.0 / sum(r["ms"] for r in cr.values() if True), "unit": "passes/s", "cores": cores, "kind": "port",
               "sample": "3 NMS settings in full + crop-resize of 200 of the 10000 boxes scaled up (the two other crop-resize rows have no CPU leg)",
               "rows": cr}
    k = rows["roi_c1_pool12_10k"]
    line = {"metric": "config5_microbench_passes_per_sec", "value": 1000.0 / total, "unit": "passes/s", "n_gpus": 1,
            "steps": max(5, args.steps), "warmup": max(3, args.warmup), "ms_per_step": total, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "l2": "256 MB flush buffer written between iterations"}, "rows": rows,
            "gpu_launches": int(launches),
            "roofline": {"kernel": "roi_kernel: crop + trilinear resize, C=1 256^3 map, 10000 boxes, pool 12^3", "bound": "hbm",
                         "achieved": k["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k["frac_of_hbm_peak"],
                         "traffic": None, "peak_source": peaks["source"] + " copy"},
            "e2e": {"value": 1000.0 / sum(r.get("e2e_ms_host_to_host", r["ms"]) for r in rows.values()), "unit": "passes/s",
                    "h2d_bytes_per_step": int(3 * n * 7 * 4), "d2h_bytes_per_step": int(4 * sum(r.get("kept", 0) for r in rows.values())),
                    "note": "NMS rows through utils.non_max_suppression with host numpy boxes in / kept indices out; crop-resize rows device-resident"},
            "cpu_baseline": cpu}
    print(json.dumps(line))
