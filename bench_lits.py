"""Part of bench.py (`--workload lits`): BASELINE.json config 3 -- the LiTS_2017 configuration of the model (P3D35 backbone,
160 / 320-channel FPN / RPN, base-32 U-Net without dropout, 3 classes, weighted CE + raw-Sobel edge loss, staged training) on a
synthetic liver-shaped input: network input 1 x 1 x 256 x 320 x 320 (what LiTS's own pad + resize produces from
512 x 512 x N scans, LiTS_2017/LiTS_main.py:117-121; that host-side preparation is off the clock).

A "step" is one train step of stage `finetune` (LiTS_2017/config.py:209-226, model.py:985-1001, 1309-1311): detector forward
(frozen) -> proposals -> 4 positive RoIs -> U-Net on 4 x (32,80,80) crops -> x2 upscale (5^3 conv) -> weighted mask CE and
raw-Sobel edge loss at (64,160,160) -> backward through the mask branch -> clip + SGD.  One JSON line."""
import json
import os
import time

import numpy as np

METRIC = "lits_volumes_per_sec_fwd_bwd"
FWD_GFLOP = 1014.4          # LiTS forward, 4 positive RoIs, of which the U-Net mask branch is 726
UNET_GFLOP = 726.0


def run_lits(args, peaks):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = ("LiTS_2017 config (P3D35, FPN 160 / RPN 320, base-32 U-Net, 3 classes): synthetic 256x320x320 input, 4 positive RoIs of "
            "(32,80,80), full train step of stage finetune (detector forward frozen, mask branch fwd + bwd + clip + SGD)")
    if args.impl != "ours":
        print(json.dumps({"impl": args.impl, "metric": METRIC, "unavailable": "the oracle port restates the heart copy of the model only; "
                          "the LiTS copy is pinned by goldens (tests/golden/lits_*.npz), not timed on the CPU"}))
        return
    import torch
    from cfun_b200 import config as Cf, model as M, ops, workload as Wk
    from cfun_b200.synth import StepInputs
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg = Cf.lits_config("finetune")
    net = M.MaskRCNN(cfg, "/tmp/_cfun_bench_lits")
    net.load_state_dict(Wk.bench_weights({k: tuple(v.shape) for k, v in net.state_dict().items()}, Wk.WEIGHT_SEED), strict=True)
    net = net.to(dev)
    H, W, D = [int(v) for v in cfg.IMAGE_SHAPE[:3]]
    anchors_np = net.anchors.cpu().numpy()
    inputs = None
    for attempt in range(32):
        seed = 3000 + attempt
        rng = np.random.default_rng(seed)
        vol = np.clip(np.round(rng.standard_normal((H, W, D), dtype=np.float32) * 300.0), -1024, 3071).astype(np.int16)
        with torch.no_grad():
            img = ops.mold_volume_i16(torch.from_numpy(vol).to(dev))
            rois = net.rpn_proposals(img, "training")[5][0]
        placed = Wk.place_label_cube(rois.cpu().numpy(), (D, H, W))
        del img, rois
        if placed is not None:
            lab = Wk.label_from_cube((D, H, W), placed[0], placed[1], seed, num_classes=cfg.NUM_CLASSES)
            inputs = StepInputs(cfg, anchors_np, None, seed, vol=vol, lab=lab)
            break
    if inputs is None:
        raise RuntimeError("no synthetic LiTS volume admits a 4-positive label placement")
    opt = net.make_optimizer(cfg.LEARNING_RATE)
    snap_p, snap_m = opt.flat_param.clone(), opt.flat_mom.clone()
    dev_inputs = [t.to(dev) for t in inputs.tensors()]

    def run_step(fn, *a):
        opt.flat_param.copy_(snap_p)
        opt.flat_mom.copy_(snap_m)
        torch.manual_seed(4321)
        return fn(opt, *a)
    run_step(net.train_step_device, *dev_inputs)
    if not args.no_graphs:
        net.enable_graphs()
    for _ in range(max(3, args.warmup)):
        run_step(net.train_step_device, *dev_inputs)
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    r0 = dict(net.graph_replays)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        last = run_step(net.train_step_device, *dev_inputs)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    launches = ops.launch_count() - l0
    for key, cnt in net.graph_replays.items():
        launches += (cnt - r0.get(key, 0)) * net.graph_kernel_counts.get(key, 0)
    for _ in range(2):
        run_step(net.train_step_from_host, inputs)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(args.steps):
        run_step(net.train_step_from_host, inputs)
    torch.cuda.synchronize()
    ms_e2e = (time.time() - t0) * 1e3
    pos, rois_n = net.last_roi_counts
    # the literal 128 -> 256 3^3 stride-2 conv of the LiTS U-Net (conv3d_c4, LiTS_2017/mask_branch.py:43) at its real extent
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x = ops.to_cl(torch.randn(4, 128, 8, 20, 20, device=dev))
    w = torch.randn(256, 128, 3, 3, 3, device=dev) * 0.02

    def med(fn, iters=7):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]
    with torch.no_grad():
        t_c4 = med(lambda: ops.conv3d(x, w, None, 2, 1))
    fl = 2.0 * 4 * 4 * 10 * 10 * 128 * 256 * 27
    step_tflop = (FWD_GFLOP + 2 * UNET_GFLOP) / 1000.0
    value = args.steps / (ms / 1000.0)
    line = {"metric": METRIC, "value": value, "unit": "volumes/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": name, "positives": pos, "rois": rois_n, "stage": "finetune", "cuda_graphs": not args.no_graphs,
                       "losses_last_step": last.cpu().numpy().tolist(),
                       "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
            "step_tflop": step_tflop, "achieved_step_tflops": value * step_tflop,
            "step_tflop_note": "forward of everything (1014.4 GFLOP, SURVEY.md 8a) + backward of the mask branch only (2 x 726): the "
                               "detector is frozen in this stage; SURVEY 8d's 3.04 TFLOP assumes a backward through everything",
            "gpu_launches": int(launches),
            "e2e": {"value": args.steps / (ms_e2e / 1000.0), "unit": "volumes/s", "h2d_bytes_per_step": int(inputs.nbytes()),
                    "d2h_bytes_per_step": 28, "ms_per_step": ms_e2e / args.steps},
            "roofline": {"kernel": "conv3d forward 3x3x3 stride 2, 128->256 @ 4 x (8,20,20) (LiTS conv3d_c4: the literal 128->256 conv), whole call "
                                   "(space-to-depth + operand packs + tcgen05 kernel)", "bound": "tensor", "achieved": fl / (t_c4 * 1e-3) / 1e12,
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / (t_c4 * 1e-3) / 1e12 / peaks["bf16_tflops"], "ms": t_c4,
                         "algorithmic_flop": fl, "traffic": None, "peak_source": peaks["source"] + " bf16 burst",
                         "note": "2.8 GFLOP on 1600 output voxels: launch- and pack-bound, not a tensor-pipe measurement"},
            "cpu_baseline": None}
    print(json.dumps(line))
