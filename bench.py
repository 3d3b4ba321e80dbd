#!/usr/bin/env python
"""Headline benchmark: 256^3 CT volumes/s, forward + backward + clip + SGD step, on N B200 GPUs (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                   (the reference's own CPU path: the oracle port, host cores)
  python bench.py --impl reference-cuda ...              (same layers through stock torch / cuDNN on the same GPU: the
                                                          library kernel-to-beat of SURVEY.md 8d)
  python bench.py --stage finetune                       (192^3 masks, 5^3 conv, Sobel edge loss; own FLOP figure)
  python bench.py --workload config5                     (BASELINE config 5: NMS + RoI crop-resize micro-benchmark)
  python bench.py --workload lits                        (BASELINE config 3: LiTS widths, 256x320x320 input)

A "step" is one pass of the hot path over one synthetic 256^3 int16 CT volume (SURVEY.md 8d recipe, HeartConfig at
256^3, 8 classes): mold -> P3D/FPN -> RPN -> proposals (sort/decode/NMS) -> detection targets -> RoI crops ->
classifier head + U-Net mask head -> six losses -> backward -> global-norm clip + SGD(momentum) -- in BOTH arms.
`value` is timed with the step's raw inputs already in HBM; `e2e` times the public call
MaskRCNN.train_step_from_host with pinned host buffers (H2D + D2H of the losses inside the timed region).
Weak scaling: every rank processes its own volume each step; the only collective is one NCCL all-reduce of the
flat gradient per step.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ct_volumes_per_sec_fwd_bwd"
PERM_SEED = 4321          # host torch.randperm draws of detection_target_layer (both arms)
DROP_SEED = 7             # Dropout3d channel masks, generated on the host and injected (both arms)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config5", "lits"])
    ap.add_argument("--image-dim", type=int, default=256)
    ap.add_argument("--stage", default="beginning", choices=["beginning", "finetune"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--profile-calls", default=None, help="write a per-library-call timing table of one eager step here")
    ap.add_argument("--kernel-table", default=None, help="write a per-kernel CUPTI table of one eager step here")
    ap.add_argument("--gap-table", default=None, help="write the device idle gaps of one step (graphs as benchmarked) here")
    ap.add_argument("--conv-algo", default="auto", choices=["auto", "simt", "tc", "tc1"])
    ap.add_argument("--no-graphs", action="store_true", help="run the heads eagerly instead of as CUDA graphs")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full` capture
    summary profiles/r02_ncu_conv_kernels.json (never a number taken in this run); None if that capture has no such row."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_conv_kernels.json")
    try:
        for row in json.load(open(p))["launches"]:
            if row.get("key") == key:
                return float(row["dram_bytes_read"]) + float(row["dram_bytes_write"])
    except Exception:
        pass
    return None


def drop_masks(n=4):
    import torch
    g = torch.Generator().manual_seed(DROP_SEED)
    return [(torch.rand(n, c, 1, 1, 1, generator=g) > 0.6).float() / 0.4 for c in (20, 40, 80, 160, 320)]


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  NVML in-process (one
    call takes microseconds, so even a 250 ms timed region gets tens of samples); `nvidia-smi` as the fallback."""

    HW_SLOWDOWN, SW_POWER_CAP, SW_THERMAL, HW_THERMAL = 0x8, 0x4, 0x20, 0x40

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []          # (sm_mhz, sm_max_mhz, set of reason names)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            names = {nm for nm, b in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                                      ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP)) if bits & b}
            self.rows.append((int(sm), int(mx), names))
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [v.strip() for v in out.split(",")]
        if len(c) >= 6 and c[0].isdigit():
            names = {nm for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[2:6])
                     if v.lower().startswith("active")}
            self.rows.append((int(c[0]), int(c[1]) if c[1].isdigit() else None, names))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:
                pass
            self.stop_flag.wait(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows if r[1]]
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores.  Imports nothing that loads libcfun_b200.so.
# ---------------------------------------------------------------------------------------------------------
def cpu_step_runner(image_dim, stage, weight_seed=None, volume_seed=None):
    """Returns (step, cores, state): step() runs ONE full train step (restore weights, forward, 6 losses, backward, clip,
    SGD with momentum / weight decay) of the CPU oracle on the synthetic volume of the shared recipe
    (cfun_b200.workload: pure numpy, no native library)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cfun_oracle as O
    from shapes import maskrcnn_shapes
    from cfun_b200 import config as Cf
    from cfun_b200 import workload as Wk
    weight_seed = Wk.WEIGHT_SEED if weight_seed is None else weight_seed
    volume_seed = Wk.VOLUME_SEED if volume_seed is None else volume_seed
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mask_pool, scales, cube = Wk.shape_params(image_dim)
    cfg = O.Cfg(image_dim=image_dim, stage=stage, mask_pool=mask_pool, anchor_scales=scales)
    pcfg = Cf.heart_config(image_dim, stage, mask_pool=mask_pool, anchor_scales=scales)
    sd = Wk.bench_weights(maskrcnn_shapes(), weight_seed)
    trainable = [k for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k and ".bn" not in k
                 and ".C1.1." not in k and "downsample.1" not in k]
    anchors = cfg.anchors().numpy()
    vol, _ = Wk.synth_volume(image_dim, volume_seed, cube)
    image = torch.from_numpy(np.ascontiguousarray(O.mold_image(vol.astype(np.float32)[..., None]).transpose((3, 2, 0, 1))[None])).float()
    # label placement from this arm's own proposals (see cfun_b200.workload.place_label_cube)
    with torch.no_grad():
        p2, p3 = O.fpn_forward(sd, image)
        lv = [O.rpn_forward(sd, p) for p in (p2, p3)]
        rois, _, _ = O.proposal_layer(torch.cat([l[1] for l in lv], 1)[0], torch.cat([l[2] for l in lv], 1)[0],
                                      cfg.anchors(), cfg.POST_NMS_ROIS_TRAINING, cfg.RPN_NMS_THRESHOLD, cfg.PRE_NMS_LIMIT,
                                      cfg.IMAGE_SHAPE, cfg.RPN_BBOX_STD_DEV)
    placed = Wk.place_label_cube(rois.numpy(), image_dim)
    if placed is None:
        raise RuntimeError("no label placement gives 4 positive RoIs for weight seed %d" % weight_seed)
    lab = Wk.label_from_cube(image_dim, placed[0], placed[1], volume_seed)
    boxes = Wk.gt_box_from_label(lab, 8)
    np.random.seed(volume_seed % (2 ** 31))
    rpn_match, rpn_bbox = Wk.build_rpn_targets(anchors, boxes[:1].astype(np.float32), pcfg)
    labt = lab.transpose((2, 0, 1))
    gt_masks = torch.from_numpy(np.stack([(labt == c) for c in range(8)]).astype(np.float32))
    args = (image, torch.from_numpy(rpn_match.astype(np.int32)), torch.from_numpy(rpn_bbox).float(), torch.arange(1, 8).int(),
            torch.from_numpy(boxes.astype(np.float32)), gt_masks)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in trainable}
    snap = {k: v.detach().clone() for k, v in leaves.items()}
    decay = [leaves[k] for k in trainable if "bn" not in k]
    nodecay = [leaves[k] for k in trainable if "bn" in k]
    groups = [{"params": decay, "weight_decay": pcfg.WEIGHT_DECAY}]
    if nodecay:
        groups.append({"params": nodecay, "weight_decay": 0.0})
    opt = torch.optim.SGD(groups, lr=pcfg.LEARNING_RATE, momentum=pcfg.LEARNING_MOMENTUM)
    sd2 = dict(sd)
    sd2.update(leaves)
    state = {"pos": None, "rois": None, "losses": None, "grad_norm": None}

    def step():
        with torch.no_grad():                 # every step is "the first step from the same checkpoint", as in the GPU arm
            for k, v in leaves.items():
                v.copy_(snap[k])
        opt.state.clear()
        opt.zero_grad(set_to_none=True)
        torch.manual_seed(PERM_SEED)
        out = O.train_forward(sd2, cfg, *args, drop=drop_masks(4))
        out["loss"].sum().backward()
        norm = torch.nn.utils.clip_grad_norm_(list(leaves.values()), 5.0)
        opt.step()
        state["pos"] = int((out["target_class_ids"] > 0).sum())
        state["rois"] = int(out["target_class_ids"].shape[0])
        state["losses"] = [float(out["loss"].sum())] + [float(l.sum()) for l in out["losses"]]
        state["grad_norm"] = float(norm)
        return out

    return step, cores, state


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cfun_b200 import workload as Wk
    step, cores, state = cpu_step_runner(args.image_dim, args.stage)
    budget = 240.0                                    # bound the whole run to a few minutes of host time
    t0 = time.time()
    step()
    t_first = time.time() - t0
    warm = max(0, min(args.warmup, int(0.25 * budget / max(t_first, 1e-3))) - 1)
    for _ in range(warm):
        step()
    k = max(1, min(args.steps, int(0.75 * budget / max(t_first, 1e-3))))
    t0 = time.time()
    for _ in range(k):
        step()
    dt = time.time() - t0
    v = k / dt
    sample = ("%d full %d^3 volume train step(s) (fwd + 6 losses + bwd + clip + SGD) of the oracle port, torch-CPU fp32, %d threads; "
              "%d warm-up step(s); %d of %d requested steps timed to bound the run" % (k, args.image_dim, cores, warm + 1, k, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "volumes/s", "n_gpus": args.gpus,
            "steps": k, "warmup": warm + 1, "ms_per_step": 1000.0 * dt / k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": Wk.workload_name(args.image_dim, args.stage), "positives": state["pos"], "rois": state["rois"],
                       "weight_seed": Wk.WEIGHT_SEED, "volume_seed": Wk.VOLUME_SEED, "losses_last_step": state["losses"],
                       "grad_norm_last_step": state["grad_norm"]},
            "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# library comparator on the same GPU: the oracle port's torch layers through cuDNN (SURVEY.md 8d "kernel to beat")
# ---------------------------------------------------------------------------------------------------------
def library_baseline(dim, stage, dev, steps=2):
    """The conv layers of one train step (P3D/FPN/RPN on the dim^3 volume, classifier head on 12 RoI crops, U-Net on
    4 x 96^3 crops + mask CE; forward + backward) as stock torch ops on the GPU: NCDHW, cudnn.benchmark on, TF32 off and
    on.  No proposal / NMS / target logic and no optimizer step, so it UNDER-counts the library's step."""
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cfun_oracle as O
    from shapes import maskrcnn_shapes
    from cfun_b200 import workload as Wk
    mask_pool, _, _ = Wk.shape_params(dim)
    sd = {k: v.to(dev) for k, v in Wk.bench_weights(maskrcnn_shapes(), Wk.WEIGHT_SEED).items()}
    leaves = [v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and v.dim() == 5]
    g = torch.Generator().manual_seed(3)
    image = torch.randn(1, 1, dim, dim, dim, generator=g).to(dev)
    pooled = torch.randn(12, 128, 12, 12, 12, generator=g).to(dev)
    crops = torch.randn(4, 1, mask_pool, mask_pool, mask_pool, generator=g).to(dev)
    side = mask_pool * (2 if stage == "finetune" else 1)
    tgt = torch.randint(0, 8, (4, side, side, side), generator=g).to(dev)
    drop = [d.to(dev) for d in drop_masks(4)]
    res = {}
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.benchmark = True
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def one():
                for v in leaves:
                    v.grad = None
                p2, p3 = O.fpn_forward(sd, image)
                lv = [O.rpn_forward(sd, p) for p in (p2, p3)]
                c_logits, _, c_bbox = O.classifier_forward(sd, pooled)
                m_logits = O.unet_forward(sd, crops, stage, drop)
                loss = sum(l[0].mean() + l[2].mean() for l in lv) + c_logits.mean() + c_bbox.mean() + F.cross_entropy(m_logits, tgt)
                loss.backward()
            one(); one()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                one()
            e.record()
            torch.cuda.synchronize()
            res["tf32_on" if tf32 else "tf32_off"] = s.elapsed_time(e) / steps
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return {"ms_per_step_fp32": res["tf32_off"], "ms_per_step_tf32": res["tf32_on"],
            "value_fp32": 1000.0 / res["tf32_off"], "value_tf32": 1000.0 / res["tf32_on"], "unit": "volumes/s",
            "covers": "conv layers of one step only (P3D/FPN/RPN at %d^3, classifier head on 12 RoIs, U-Net on 4x%d^3, mask CE; "
                      "fwd+bwd) through torch %s / cuDNN %s, NCDHW, cudnn.benchmark=True; no proposal/NMS/target logic, no "
                      "optimizer step" % (dim, mask_pool, torch.__version__, torch.backends.cudnn.version())}


def run_reference_cuda(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from cfun_b200 import workload as Wk
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lb = library_baseline(args.image_dim, args.stage, dev, steps=max(1, args.steps))
    line = {"impl": "reference-cuda", "metric": METRIC, "value": lb["value_fp32"], "unit": "volumes/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 2, "ms_per_step": lb["ms_per_step_fp32"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": Wk.workload_name(args.image_dim, args.stage) + " -- LAYERS ONLY, see covers"}, "gpu_library_baseline": lb}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=5, flush=None, kernel_only=False):
    """median CUDA-event time of fn() on the current stream; kernel_only: the library's own bracket around the main kernel
    of the LAST library call fn() makes (cfun_kernel_timing)"""
    import torch
    from cfun_b200 import ops
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(ops.last_kernel_ms() if kernel_only else s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def build_gpu_case(dim, stage, dev, rank=0, conv_algo="auto"):
    """Model with the shared weights + one synthetic volume whose label placement gives 4 positives / 12 RoIs (off the
    clock).  Used by bench.py and by tests/test_gpu_model.py::test_train_step_256_matches_cpu_oracle."""
    import torch
    from cfun_b200 import model as M, config as Cf, ops
    from cfun_b200 import workload as Wk
    from cfun_b200.synth import StepInputs
    ops.set_conv_algo({"auto": ops.ALGO_AUTO, "simt": ops.ALGO_SIMT, "tc": ops.ALGO_TC, "tc1": ops.ALGO_TC1}[conv_algo])
    mask_pool, scales, cube = Wk.shape_params(dim)
    cfg = Cf.heart_config(dim, stage, mask_pool=mask_pool, anchor_scales=scales)
    net = M.MaskRCNN(cfg, "/tmp/_cfun_bench")
    net.load_state_dict(Wk.bench_weights({k: tuple(v.shape) for k, v in net.state_dict().items()}, Wk.WEIGHT_SEED), strict=True)
    net = net.to(dev)
    anchors_np = net.anchors.cpu().numpy()
    for attempt in range(32):       # a volume whose untrained proposals admit a 4-positive label placement
        vol_seed = Wk.VOLUME_SEED + rank * 10007 + attempt
        vol, _ = Wk.synth_volume(dim, vol_seed, cube)
        with torch.no_grad():
            img = ops.mold_volume_i16(torch.from_numpy(vol).to(dev))
            rois = net.rpn_proposals(img, "training")[5][0]
        placed = Wk.place_label_cube(rois.cpu().numpy(), dim)
        del img, rois
        if placed is not None:
            lab = Wk.label_from_cube(dim, placed[0], placed[1], vol_seed)
            return net, cfg, StepInputs(cfg, anchors_np, dim, vol_seed, cube, vol=vol, lab=lab), vol_seed
    raise RuntimeError("no synthetic volume admits a 4-positive label placement (rank %d)" % rank)


def kernel_rooflines(dev, peaks, stage):
    """Live rooflines of the dominant kernels: CUDA events around the main kernel itself (cfun_kernel_timing), L2 flushed
    between iterations.  tensor: algorithmic FLOPs / measured bf16 burst peak; hbm: algorithmic bytes / measured copy peak."""
    import torch
    from cfun_b200 import ops
    from cfun_b200.layers import Conv3d
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)
    out = {}
    ops.kernel_timing(True)
    try:
        def conv_case(N, Ci, S, Co, name, module, key_prefix):
            conv = Conv3d(Ci, Co, 3, padding=1, bias=False).to(dev)
            x = ops.to_cl(torch.randn(N, Ci, S, S, S, device=dev)).requires_grad_(True)
            fl = 2.0 * N * S ** 3 * Ci * Co * 27
            y = conv(x)
            dy = ops.to_cl(torch.randn_like(y))
            rows = {}
            with torch.no_grad():
                t_all = time_kernel(lambda: conv(x), flush=flush)
                rows["fwd"] = (time_kernel(lambda: conv(x), flush=flush, kernel_only=True), t_all)

            def bwd():
                x.grad = None
                conv.weight.grad = None
                conv(x).backward(dy)
            # fused backward: dY pack -> data-gradient kernel -> weight-gradient kernel (the last one is what the bracket holds)
            rows["wgrad"] = (time_kernel(bwd, flush=flush, kernel_only=True), None)
            res = {}
            for ps, (t, t_call) in rows.items():
                ach = fl / (t * 1e-3) / 1e12
                res[ps] = {"kernel": "%s %s 3x3x3 %d->%d @ %dx%d^3 (%s), main kernel only" % (
                               {"fwd": "conv3d forward", "wgrad": "conv3d weight gradient"}[ps], module, Ci, Co, N, S, name),
                           "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                           "frac": ach / peaks["bf16_tflops"], "ms": t, "algorithmic_flop": fl,
                           "traffic": ncu_traffic("%s_%s" % (key_prefix, ps)),
                           "traffic_unit": "bytes per launch (dram read + write, profiles/r02_ncu_conv_kernels.json)",
                           "peak_source": peaks["source"] + " bf16 burst"}
                if t_call is not None:
                    res[ps]["ms_with_operand_packs"] = t_call
            del x, y, dy, conv
            return res
        u = conv_case(4, 40, 96, 40, "mask_branch conv_norm_lrelu_l4.0", "conv_tc_halo_kernel / conv_tc_wgrad_ds_kernel", "unet40")
        h = conv_case(1, 128, 32, 256, "RPN.conv_shared, north-star headline conv", "conv_tc_hx_kernel / conv_tc_wgrad_ds_kernel", "rpn")
        note = ("parity mode issues 3 bf16 MMAs per product (split-bf16, DESIGN.md 4): frac is capped at 1/3; the ncu "
                "tensor-pipe-active figures are in profiles/r02_ncu_conv_kernels.json")
        out["roofline"] = dict(u["wgrad"], note=note + "; this is the kernel with the largest share of the step "
                               "(profiles/r02_step_kernel_table.txt)")
        out["roofline_fwd"] = dict(u["fwd"], note=note)
        out["headline_conv"] = {"fwd": h["fwd"], "wgrad": h["wgrad"]}
    finally:
        ops.kernel_timing(False)
    # dominant element-wise pass: backward of InstanceNorm + LeakyReLU over a 40-channel 4 x 96^3 tensor (reads x and dy, writes dx)
    x = ops.to_cl(torch.randn(4, 40, 96, 96, 96, device=dev)).requires_grad_(True)
    y = ops.instnorm_lrelu(x)
    dy = ops.to_cl(torch.randn_like(y))

    def in_bwd():
        x.grad = None
        ops.instnorm_lrelu(x).backward(dy)
    with torch.no_grad():
        t_f = time_kernel(lambda: ops.instnorm_lrelu(x), flush=flush)
    t_fb = time_kernel(in_bwd, flush=flush)
    nb = x.numel() * 4.0
    # forward: stats pass reads x, apply pass reads x and writes y (3 tensor passes); backward: pass 1 reads x, dy and writes a partial
    # dx, pass 2 reads x and the partial and writes dx (5 tensor passes, 3 algorithmic: read x, read dy, write dx)
    gbs = 3 * nb / ((t_fb - t_f) * 1e-3) / 1e9
    out["roofline_hbm"] = {"kernel": "InstanceNorm+LeakyReLU backward (affine_act_bwd + in_bwd_apply) over 40 ch x 4 x 96^3, whole op",
                           "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                           "ms": t_fb - t_f, "algorithmic_bytes": 3 * nb, "traffic": None, "peak_source": peaks["source"] + " copy"}
    del x, y, dy
    if stage == "finetune":
        P, M = 4, 192
        pred = ops.to_cl(torch.softmax(torch.randn(P, 8, M, M, M, device=dev), 1)).requires_grad_(True)
        tgt = torch.randint(0, 8, (P, M, M, M), device=dev)
        with torch.no_grad():
            t_f = time_kernel(lambda: ops.sobel_edge_loss(pred, tgt), flush=flush)

        def fb():
            pred.grad = None
            ops.sobel_edge_loss(pred, tgt).backward()
        t_fb = time_kernel(fb, flush=flush)
        alg = 2.0 * P * 7 * M ** 3 * 4                  # predicted and target planes of the 7 foreground classes, read once
        out["roofline_sobel"] = {"kernel": "3-D Sobel edge loss forward (sobel_pass1/2) at 4 x 7 x 192^3", "bound": "hbm",
                                 "achieved": alg / (t_f * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": alg / (t_f * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": t_f, "ms_fwd_bwd": t_fb,
                                 "algorithmic_bytes": alg, "traffic": None, "peak_source": peaks["source"] + " copy"}
        del pred, tgt
    del flush
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cfun_b200 import ops
    from cfun_b200 import workload as Wk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs CUDA; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    dim = args.image_dim
    # weights: the reference's own initialisation (MaskRCNN.initialize_weights) from a per-tensor seeded recipe shared with
    # the CPU arm; label cube placed where the untrained detector's proposals cluster so that the sampled RoI set is the
    # 4 positives / 12 RoIs the FLOP figure is quoted for
    net, cfg, inputs, vol_seed = build_gpu_case(dim, args.stage, dev, rank, args.conv_algo)
    pool = [inputs]
    net.mask.modified_u_net.injected_drop = [d.to(dev) for d in drop_masks(4)]   # same host-generated Dropout3d masks as the CPU arm
    opt = net.make_optimizer(cfg.LEARNING_RATE)
    if world > 1:   # identical replicas: broadcast rank 0's parameters once
        dist.broadcast(opt.flat_param, src=0)
        opt.overlap_allreduce()      # classifier.conv1's 113 MB gradient slice is reduced on a side stream while backward continues
    # Every timed step is "the first step from the same checkpoint": with random-init weights all RPN scores sit within
    # ~1e-3 of 0.5, so ANY parameter update reshuffles the top-1000 anchors and the sampled RoI set (and with it 92 % of
    # the FLOPs) would drift from step to step.  Parameters and momentum are therefore restored from a device snapshot at
    # the start of each step -- 2 x 165 MB of extra D2D copies INSIDE the timed region, no work removed: the optimizer
    # step still runs in full.
    snap_p, snap_m = opt.flat_param.clone(), opt.flat_mom.clone()

    def run_step(fn, *a):
        opt.flat_param.copy_(snap_p)
        opt.flat_mom.copy_(snap_m)
        torch.manual_seed(PERM_SEED)       # same host randperm draws -> same sampled RoIs (and the same as the CPU arm)
        return fn(opt, *a)
    dev_inputs = [[t.to(dev) for t in p.tensors()] for p in pool]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------
    run_step(net.train_step_device, *dev_inputs[0])      # eager step: sizes the conv workspace, fills caches
    if not args.no_graphs:
        net.enable_graphs()                              # heads replay as CUDA graphs (captured in the first warm-up step)
    for i in range(args.warmup):
        run_step(net.train_step_device, *dev_inputs[i % len(dev_inputs)])
    barrier()
    replays0 = dict(net.graph_replays)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    last = None
    roi_counts = []
    for i in range(args.steps):
        last = run_step(net.train_step_device, *dev_inputs[i % len(dev_inputs)])
        roi_counts.append(list(net.last_roi_counts))
    e.record()
    barrier()
    clocks = sampler.summary()          # samples cover exactly the timed region (GPU busy from s to e)
    ms = s.elapsed_time(e)
    launches = ops.launch_count() - l0
    for key, cnt in net.graph_replays.items():          # kernels inside replayed graphs are not seen by the launch hook
        launches += (cnt - replays0.get(key, 0)) * net.graph_kernel_counts.get(key, 0)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    losses = last.cpu().numpy().tolist()
    pos, rois = net.last_roi_counts

    # ---- end-to-end timing through the public host call ---------------------------------------------------
    for i in range(min(args.warmup, 2)):
        run_step(net.train_step_from_host, pool[i % len(pool)])
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    pending, host_losses = None, []
    for i in range(args.steps):      # every step: H2D of its inputs, the step, D2H of its 7 losses -- read on the host one step later
        nxt = run_step(net.train_step_from_host, pool[i % len(pool)], False)
        if pending is not None:
            host_losses.append(pending.result().tolist())
        pending = nxt
    host_losses.append(pending.result().tolist())
    e2.record()
    assert len(host_losses) == args.steps and all(len(r) == 7 for r in host_losses)
    barrier()
    t2 = torch.tensor([s2.elapsed_time(e2)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2.item())

    if args.profile_calls and rank == 0:   # off the clock: one eager step with per-library-call CUDA-event timing
        net.enable_graphs(False)
        ops.profile_start()
        run_step(net.train_step_device, *dev_inputs[0])
        rows = ops.profile_stop()
        agg = {}
        for name, tag, tt in rows:
            a = agg.setdefault((name, tag), [0, 0.0])
            a[0] += 1
            a[1] += tt
        tot = sum(tt for _, _, tt in rows)
        with open(args.profile_calls, "w") as f:
            f.write("# library calls of one eager train step, CUDA-event time per call (includes inter-call gaps on the stream)\n")
            f.write("# total %.2f ms over %d calls\n" % (tot, len(rows)))
            for (name, tag), (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%8.3f ms %5.1f%% x%-4d %s %s\n" % (tt, 100 * tt / tot, n, name, tag))
    if args.kernel_table and rank == 0:    # off the clock: one eager step under CUPTI (torch.profiler), device time per kernel
        from torch.profiler import profile, ProfilerActivity
        net.enable_graphs(False)
        run_step(net.train_step_device, *dev_inputs[0])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run_step(net.train_step_device, *dev_inputs[0])
            torch.cuda.synchronize()
        agg = {}
        for ev in prof.events():
            if ev.device_type is not None and str(ev.device_type).endswith("CUDA"):
                a = agg.setdefault(ev.name, [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time_total / 1000.0 if hasattr(ev, "device_time_total") else ev.cuda_time_total / 1000.0
        tot = sum(v[1] for v in agg.values())
        with open(args.kernel_table, "w") as f:
            f.write("# one eager %d^3 train step, stage %s (%d positive / %d RoIs), device time per kernel from CUPTI (torch.profiler)\n" % (dim, args.stage, pos, rois))
            f.write("# total kernel time %.2f ms over %d launches\n" % (tot, sum(v[0] for v in agg.values())))
            for name, (n, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%9.3f ms %5.1f%% %5d  %s\n" % (tt, 100 * tt / tot, n, name[:140]))

    if args.gap_table and rank == 0:       # off the clock: where the device idles inside one step as benchmarked
        from torch.profiler import profile, ProfilerActivity
        net.enable_graphs(True)
        for _ in range(2):
            run_step(net.train_step_device, *dev_inputs[0])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):      # back to back as in the timed loop: the host runs ahead of the device across steps
                run_step(net.train_step_device, *dev_inputs[0])
            torch.cuda.synchronize()
        evs = sorted((ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
                     if ev.device_type is not None and str(ev.device_type).endswith("CUDA"))
        marks = [e[0] for e in evs if "mold_transpose_kernel" in e[2]]      # one per step
        if len(marks) >= 4:
            evs = [e for e in evs if marks[2] <= e[0] < marks[3]]          # the third step, steady state
        busy_end, gaps, busy = evs[0][0], [], 0.0
        prev_name = "-"
        for st_, en_, name in evs:
            if st_ > busy_end:
                gaps.append((st_ - busy_end, busy_end - evs[0][0], prev_name, name))
            if en_ > busy_end:
                busy += en_ - max(st_, busy_end)
                busy_end, prev_name = en_, name
        span = busy_end - evs[0][0]
        with open(args.gap_table, "w") as f:
            f.write("# one %d^3 train step as benchmarked (CUDA graphs on, third of four back-to-back steps), CUPTI timeline: first kernel start -> last kernel end\n" % dim)
            f.write("# span %.3f ms, device busy %.3f ms, idle %.3f ms in %d gaps (%d launches)\n"
                    % (span / 1e3, busy / 1e3, (span - busy) / 1e3, len(gaps), len(evs)))
            f.write("# idle by size: " + ", ".join("%s %.3f ms" % (lbl, sum(g[0] for g in gaps if lo_ <= g[0] < hi_) / 1e3)
                                                  for lbl, lo_, hi_ in (("<5us", 0, 5), ("5-20us", 5, 20), ("20-100us", 20, 100),
                                                                        (">=100us", 100, 1e12))) + "\n")
            f.write("# gap_us  at_ms  after kernel -> before kernel\n")
            for g, at, a, b in sorted(gaps, reverse=True)[:60]:
                f.write("%8.1f %7.3f  %s -> %s\n" % (g, at / 1e3, a[:70], b[:70]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * args.steps / (ms / 1000.0)
    e2e_value = world * args.steps / (ms_e2e / 1000.0)
    peaks = measured_peaks()
    step_tflop = Wk.STEP_TFLOP[args.stage]
    roofs = kernel_rooflines(dev, peaks, args.stage) if dim >= 256 else {}

    cpu = None
    loss_check = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            step, cores, st = cpu_step_runner(dim, args.stage)
            step()                                   # warm-up (thread pool, oneDNN primitive caches)
            t0 = time.time()
            step()
            dt = time.time() - t0
            cpu = {"value": 1.0 / dt, "unit": "volumes/s", "cores": cores, "kind": "port",
                   "sample": "1 full %d^3 volume train step (fwd + 6 losses + bwd + clip + SGD) of the CPU oracle port, torch-CPU fp32, "
                             "after 1 warm-up step, %d positives / %d RoIs" % (dim, st["pos"], st["rois"]),
                   "losses": st["losses"]}
            if (st["pos"], st["rois"]) == (pos, rois):
                d = [abs(a - b) / max(abs(b), 1e-6) for a, b in zip(losses, st["losses"]) if abs(b) > 0]
                loss_check = {"max_rel_diff_vs_cpu_oracle": max(d), "compared": "total + six losses of the same step, same inputs / weights / draws"}
        except Exception as ex:   # never let the baseline leg break the measurement
            cpu = {"value": None, "unit": "volumes/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}
    lib = None
    if world == 1 and not args.no_library_baseline and dim >= 128:
        try:
            lib = library_baseline(dim, args.stage, dev)
        except Exception as ex:
            lib = {"failed": repr(ex)}

    h2d = pool[0].nbytes()
    line = {
        "metric": METRIC, "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": Wk.workload_name(dim, args.stage), "parallelism": "dp%d" % world, "positives": pos, "rois": rois,
                   "roi_counts_per_timed_step": roi_counts, "weight_seed": Wk.WEIGHT_SEED, "volume_seed": vol_seed,
                   "conv_algo": args.conv_algo, "cuda_graphs": (not args.no_graphs),
                   "l2": "per-step working set (~10 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "losses_last_step": losses},
        "step_tflop": step_tflop, "achieved_step_tflops": value * step_tflop / max(world, 1),
        "step_frac_of_bf16_sustained": value * step_tflop / max(world, 1) / peaks["bf16_tflops_sustained"],
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 7 * 4,
                "ms_per_step": ms_e2e / args.steps,
                "readback": "the 7 losses of every step are copied to pinned host memory and read inside the timed region, "
                            "step i's while step i+1 runs (MaskRCNN.train_step_from_host(wait=False))"},
        "roofline": roofs.get("roofline"), "roofline_fwd": roofs.get("roofline_fwd"), "roofline_hbm": roofs.get("roofline_hbm"),
        "headline_conv": roofs.get("headline_conv"), "cpu_baseline": cpu, "loss_check": loss_check, "gpu_library_baseline": lib,
    }
    if "roofline_sobel" in roofs:
        line["roofline_sobel"] = roofs["roofline_sobel"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.workload == "config5":
        from bench_config5 import run_config5
        run_config5(a, measured_peaks())
    elif a.workload == "lits":
        from bench_lits import run_lits
        run_lits(a, measured_peaks())
    elif a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-cuda":
        run_reference_cuda(a)
    else:
        run_ours(a)
