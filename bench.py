#!/usr/bin/env python
"""Headline benchmark: 256^3 CT volumes/s, forward + backward + clip + SGD step, on N B200 GPUs (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                   (the reference's own CPU path: the oracle port, host cores)

A "step" is one pass of the hot path over one synthetic 256^3 int16 CT volume (SURVEY.md 8d recipe, HeartConfig at
256^3, 8 classes, stage 'beginning'): mold -> P3D/FPN -> RPN -> proposals (sort/decode/NMS) -> detection targets ->
RoI crops -> classifier head + U-Net mask head -> six losses -> backward -> global-norm clip + SGD(momentum).
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

FWD_GFLOP_PER_VOLUME = 1314.6        # SURVEY.md 8d / BASELINE.md: H256, 4 positive / 12 RoIs, 'beginning'
STEP_TFLOP_PER_VOLUME = 3.94


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--image-dim", type=int, default=256)
    ap.add_argument("--stage", default="beginning")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-calls", default=None, help="write a per-library-call timing table of one eager step here")
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


def bench_weights(shapes, seed):
    """MaskRCNN.initialize_weights recipe (reference model.py:1306-1319: xavier_uniform conv weights, zero conv bias,
    N(0, 0.01) linear weights, BN at identity) with a per-tensor generator keyed by (seed, name), so the GPU arm and the
    CPU reference arm can build bit-identical weights from a seed alone."""
    import math
    import zlib
    import torch
    sd = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        g = torch.Generator().manual_seed((seed * 7919 + zlib.crc32(k.encode())) % (2 ** 31 - 1))
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
        elif k.endswith("running_var") or (k.endswith(".weight") and len(shp) == 1):
            sd[k] = torch.ones(shp)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            sd[k] = torch.zeros(shp)
        elif len(shp) == 5:
            rf = shp[2] * shp[3] * shp[4]
            bound = math.sqrt(6.0 / (shp[1] * rf + shp[0] * rf))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.01
    return sd


WEIGHT_SEED = 2     # chosen so that the synthetic volume yields 4 positive / 12 sampled RoIs


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
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_step_runner(image_dim, stage, weight_seed=WEIGHT_SEED):
    """Returns (fn, cores): fn() runs ONE full train step (forward, 6 losses, backward, clip) of the CPU oracle on a
    synthetic volume with the same recipe as the GPU arm."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cfun_oracle as O
    from shapes import maskrcnn_shapes
    from cfun_b200 import config as Cf
    from cfun_b200.synth import synth_volume, gt_box_from_label, place_label_cube, label_from_cube
    import cfun_b200.model as M                      # host-side target builder only (numpy)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mask_pool = 96 if image_dim >= 128 else 32
    scales = (64, 128) if image_dim >= 256 else ((32, 64) if image_dim >= 128 else (16, 32))
    cube = 70 * image_dim // 256
    cfg = O.Cfg(image_dim=image_dim, stage=stage, mask_pool=mask_pool, anchor_scales=scales)
    pcfg = Cf.heart_config(image_dim, stage, mask_pool=mask_pool, anchor_scales=scales)
    sd = bench_weights(maskrcnn_shapes(), weight_seed)
    trainable = [k for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k and ".bn" not in k
                 and ".C1.1." not in k and "downsample.1" not in k]
    anchors = cfg.anchors().numpy()
    vol, _ = synth_volume(image_dim, 1000, cube)
    image = torch.from_numpy(np.ascontiguousarray(O.mold_image(vol.astype(np.float32)[..., None]).transpose((3, 2, 0, 1))[None])).float()
    # label placement from this arm's own proposals (see cfun_b200.synth.place_label_cube)
    with torch.no_grad():
        p2, p3 = O.fpn_forward(sd, image)
        lv = [O.rpn_forward(sd, p) for p in (p2, p3)]
        rois, _, _ = O.proposal_layer(torch.cat([l[1] for l in lv], 1)[0], torch.cat([l[2] for l in lv], 1)[0],
                                      cfg.anchors(), cfg.POST_NMS_ROIS_TRAINING, cfg.RPN_NMS_THRESHOLD, cfg.PRE_NMS_LIMIT,
                                      cfg.IMAGE_SHAPE, cfg.RPN_BBOX_STD_DEV)
    placed = place_label_cube(rois.numpy(), image_dim)
    if placed is None:
        raise RuntimeError("no label placement gives 4 positive RoIs for weight seed %d" % weight_seed)
    lab = label_from_cube(image_dim, placed[0], placed[1], 1000)
    boxes = gt_box_from_label(lab, 8)
    np.random.seed(1000)
    rpn_match, rpn_bbox = M.build_rpn_targets(anchors, boxes[:1].astype(np.float32), pcfg)
    labt = lab.transpose((2, 0, 1))
    gt_masks = torch.from_numpy(np.stack([(labt == c) for c in range(8)]).astype(np.float32))
    args = (image, torch.from_numpy(rpn_match.astype(np.int32)), torch.from_numpy(rpn_bbox).float(), torch.arange(1, 8).int(),
            torch.from_numpy(boxes.astype(np.float32)), gt_masks)
    state = {"pos": None}

    def step():
        g = torch.Generator().manual_seed(7)
        torch.manual_seed(77)
        leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in trainable}
        sd2 = dict(sd)
        sd2.update(leaves)
        # Dropout3d draws for up to 4 positives; the oracle slices them to the actual positive count
        drop = [(torch.rand(4, c, 1, 1, 1, generator=g) > 0.6).float() / 0.4 for c in (20, 40, 80, 160, 320)]
        out = O.train_forward(sd2, cfg, *args, drop=drop)
        out["loss"].sum().backward()
        torch.nn.utils.clip_grad_norm_(list(leaves.values()), 5.0)
        state["pos"] = int((out["target_class_ids"] > 0).sum())
        state["rois"] = int(out["target_class_ids"].shape[0])
        return out

    return step, cores, state


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores, state = cpu_step_runner(args.image_dim, args.stage)
    t0 = time.time()
    warm = min(args.warmup, 1)
    for _ in range(warm):
        step()
    t_first = time.time() - t0 if warm else None
    budget = 240.0
    k = args.steps
    if t_first:
        k = max(1, min(args.steps, int(budget / max(t_first, 1e-3))))
    t0 = time.time()
    for _ in range(k):
        step()
    dt = time.time() - t0
    v = k / dt
    sample = "%d full %d^3 volume train step(s) (fwd+6 losses+bwd+clip) of the oracle port, torch-CPU fp32, %d threads; %d of %d requested steps timed to bound the run" % (
        k, args.image_dim, cores, k, args.steps)
    line = {"impl": "reference", "metric": "ct_volumes_per_sec_fwd_bwd", "value": v, "unit": "volumes/s", "n_gpus": args.gpus,
            "steps": k, "warmup": warm, "ms_per_step": 1000.0 * dt / k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MM-WHS-shape %d^3 synthetic CT, 8-class heart, full train step, stage %s" % (args.image_dim, args.stage),
                       "positives": state.get("pos"), "rois": state.get("rois")},
            "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=5, flush=None):
    import torch
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
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cfun_b200 import model as M, config as Cf, ops
    from cfun_b200.synth import StepInputs

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
    algo = {"auto": ops.ALGO_AUTO, "simt": ops.ALGO_SIMT, "tc": ops.ALGO_TC, "tc1": ops.ALGO_TC1}[args.conv_algo]
    ops.set_conv_algo(algo)

    dim = args.image_dim
    mask_pool = 96 if dim >= 128 else 32
    scales = (64, 128) if dim >= 256 else ((32, 64) if dim >= 128 else (16, 32))
    cube = 70 * dim // 256
    cfg = Cf.heart_config(dim, args.stage, mask_pool=mask_pool, anchor_scales=scales)

    # weights: the reference's own initialisation (MaskRCNN.initialize_weights, bit-identical RNG consumption) from a
    # per-tensor seeded recipe shared with the CPU arm; label cube placed where the untrained detector's proposals
    # cluster so that the sampled RoI set is the 4 positives / 12 RoIs the FLOP figure is quoted for
    from cfun_b200.synth import synth_volume, place_label_cube, label_from_cube
    net = M.MaskRCNN(cfg, "/tmp/_cfun_bench")
    net.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in net.state_dict().items()}, WEIGHT_SEED), strict=True)
    net = net.to(dev)
    anchors_np = net.anchors.cpu().numpy()
    pool = []
    vol_seed = None
    for attempt in range(32):       # a volume whose untrained proposals admit a 4-positive label placement (off the clock)
        vol_seed = 1000 + rank * 10007 + attempt
        vol, _ = synth_volume(dim, vol_seed, cube)
        with torch.no_grad():
            img = ops.mold_volume_i16(torch.from_numpy(vol).to(dev))
            rois = net.rpn_proposals(img, "training")[5][0]
        placed = place_label_cube(rois.cpu().numpy(), dim)
        del img, rois
        if placed is not None:
            lab = label_from_cube(dim, placed[0], placed[1], vol_seed)
            pool.append(StepInputs(cfg, anchors_np, dim, vol_seed, cube, vol=vol, lab=lab))
            break
    if not pool:
        raise RuntimeError("no synthetic volume admits a 4-positive label placement (rank %d)" % rank)
    weight_seed = WEIGHT_SEED
    opt = net.make_optimizer(cfg.LEARNING_RATE)
    if world > 1:   # identical replicas: broadcast rank 0's parameters once
        dist.broadcast(opt.flat_param, src=0)
    # Every timed step is "the first step from the same checkpoint": with random-init weights all RPN scores sit within
    # ~1e-3 of 0.5, so ANY parameter update reshuffles the top-1000 anchors and the sampled RoI set (and with it 92 % of
    # the FLOPs) would drift from step to step.  Parameters and momentum are therefore restored from a device snapshot at
    # the start of each step -- 2 x 165 MB of extra D2D copies INSIDE the timed region, no work removed: the optimizer
    # step still runs in full.
    snap_p, snap_m = opt.flat_param.clone(), opt.flat_mom.clone()

    def run_step(fn, *a):
        opt.flat_param.copy_(snap_p)
        opt.flat_mom.copy_(snap_m)
        torch.manual_seed(4321 + rank)       # same host randperm draws -> same sampled RoIs
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
    for i in range(args.steps):
        out = run_step(net.train_step_from_host, pool[i % len(pool)])
    e2.record()
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
        for name, tag, t in rows:
            a = agg.setdefault((name, tag), [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(t for _, _, t in rows)
        with open(args.profile_calls, "w") as f:
            f.write("# library calls of one eager train step, CUDA-event time per call (includes inter-call gaps on the stream)\n")
            f.write("# total %.2f ms over %d calls\n" % (tot, len(rows)))
            for (name, tag), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%8.3f ms %5.1f%% x%-4d %s %s\n" % (t, 100 * t / tot, n, name, tag))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * args.steps / (ms / 1000.0)
    e2e_value = world * args.steps / (ms_e2e / 1000.0)
    peaks = measured_peaks()

    # ---- roofline of the dominant kernel, measured live with CUDA events on our stream ------------------------
    from cfun_b200.layers import Conv3d
    roof = kern = None
    if dim >= 256:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)
        conv = Conv3d(40, 40, 3, padding=1, bias=False).to(dev)
        x = ops.to_cl(torch.randn(4, 40, 96, 96, 96, device=dev))
        with torch.no_grad():
            t_dom = time_kernel(lambda: conv(x), flush=flush)
        fl = 2.0 * 4 * 96 ** 3 * 40 * 40 * 27
        ach = fl / (t_dom * 1e-3) / 1e12
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of conv_tc_halo_kernel for this launch, from the committed
        # ncu --set full capture profiles/r01_ncu_unet_halo_fwd_dgrad_ds_wgrad.json (694.1 MB read: the two split-bf16
        # activation packs once; 532.6 MB written: the fp32 output once) -- equal to the algorithmic bytes, no re-reads
        roof = {"kernel": "conv3d fwd 3x3x3 40->40 @ 4x96^3 (mask_branch conv_norm_lrelu_l4.0, largest FLOP share of the step; "
                          "pack_act_gp + pack_w_halo + conv_tc_halo_kernel)",
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": 1226.7e6, "traffic_unit": "bytes per launch (ncu dram read + write)",
                "peak_source": peaks["source"] + " bf16 burst",
                "ms": t_dom, "algorithmic_flop": fl,
                "note": "parity mode issues 3 bf16 MMAs per product (split-bf16, DESIGN.md 4): frac is capped at 1/3; ncu "
                        "tensor-pipe active 57 % for this kernel, 80 % for the 128->256 headline conv"}
        conv2 = Conv3d(128, 256, 3, padding=1).to(dev)
        x2 = ops.to_cl(torch.randn(1, 128, 32, 32, 32, device=dev))
        with torch.no_grad():
            t_h = time_kernel(lambda: conv2(x2), flush=flush)
        fl2 = 2.0 * 32 ** 3 * 128 * 256 * 27
        kern = {"kernel": "conv3d fwd 3x3x3 128->256 @ 32^3 (RPN.conv_shared, north-star headline conv)", "ms": t_h,
                "achieved_tflops": fl2 / (t_h * 1e-3) / 1e12, "frac_of_bf16_peak": fl2 / (t_h * 1e-3) / 1e12 / peaks["bf16_tflops"]}
        del flush, x, x2

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            step, cores, st = cpu_step_runner(dim, args.stage)
            t0 = time.time()
            step()
            dt = time.time() - t0
            cpu = {"value": 1.0 / dt, "unit": "volumes/s", "cores": cores, "kind": "port",
                   "sample": "1 full %d^3 volume train step (fwd+6 losses+bwd+clip) of the CPU oracle port, torch-CPU fp32, cold "
                             "(no warm-up), %d positives / %d RoIs" % (dim, st.get("pos"), st.get("rois"))}
        except Exception as ex:   # never let the baseline leg break the measurement
            cpu = {"value": None, "unit": "volumes/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}

    h2d = pool[0].nbytes()
    line = {
        "metric": "ct_volumes_per_sec_fwd_bwd", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MM-WHS-shape %d^3 synthetic int16 CT, 8-class heart, full train step (fwd+bwd+clip+SGD), stage %s, 1 volume/GPU/step" % (dim, args.stage),
                   "parallelism": "dp%d" % world, "positives": pos, "rois": rois, "roi_counts_per_timed_step": roi_counts, "weight_seed": weight_seed, "volume_seed": vol_seed,
                   "conv_algo": args.conv_algo, "cuda_graphs": (not args.no_graphs), "l2": "per-step working set (~10 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "losses_last_step": losses},
        "step_tflop": STEP_TFLOP_PER_VOLUME, "achieved_step_tflops": value * STEP_TFLOP_PER_VOLUME / max(world, 1),
        "step_frac_of_bf16_sustained": value * STEP_TFLOP_PER_VOLUME / max(world, 1) / peaks["bf16_tflops_sustained"],
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 7 * 4,
                "ms_per_step": ms_e2e / args.steps},
        "roofline": roof, "headline_conv": kern, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
