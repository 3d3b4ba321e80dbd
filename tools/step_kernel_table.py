"""Per-kernel device-time table of one 256^3 train step (CUPTI through torch.profiler; GPU box only).

  python tools/step_kernel_table.py > profiles/rNN_step_kernel_table.txt
Same model / inputs / weights as bench.py; one eager step (no CUDA graphs) after two warm-up steps.  Unlike the ncu launch
list the kernels run back to back at full speed with warm caches, so these are the real in-step durations."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cfun_b200 import model as M, config as Cf, ops
from cfun_b200.synth import StepInputs, synth_volume, place_label_cube, label_from_cube

dev = torch.device("cuda", 0)
dim, cube = 256, 70
cfg = Cf.heart_config(dim, "beginning", mask_pool=96, anchor_scales=(64, 128))
net = M.MaskRCNN(cfg, "/tmp/_cfun_bench")
net.load_state_dict(bench.bench_weights({k: tuple(v.shape) for k, v in net.state_dict().items()}, bench.WEIGHT_SEED), strict=True)
net = net.to(dev)
anchors_np = net.anchors.cpu().numpy()
for attempt in range(32):
    seed = 1000 + attempt
    vol, _ = synth_volume(dim, seed, cube)
    with torch.no_grad():
        img = ops.mold_volume_i16(torch.from_numpy(vol).to(dev))
        rois = net.rpn_proposals(img, "training")[5][0]
    placed = place_label_cube(rois.cpu().numpy(), dim)
    if placed is not None:
        break
lab = label_from_cube(dim, placed[0], placed[1], seed)
inp = [t.to(dev) for t in StepInputs(cfg, anchors_np, dim, seed, cube, vol=vol, lab=lab).tensors()]
opt = net.make_optimizer(cfg.LEARNING_RATE)
snap_p, snap_m = opt.flat_param.clone(), opt.flat_mom.clone()


def step():
    opt.flat_param.copy_(snap_p)
    opt.flat_mom.copy_(snap_m)
    torch.manual_seed(4321)
    return net.train_step_device(opt, *inp)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(ev.device_time_total / 1e3, ev.count, ev.key) for ev in prof.key_averages() if ev.device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("# one eager 256^3 train step (4 positive / %d RoIs), device time per kernel from CUPTI (torch.profiler)" % net.last_roi_counts[1])
print("# total kernel time %.2f ms over %d launches" % (tot, sum(r[1] for r in rows)))
for t, n, k in rows[:70]:
    print("%9.3f ms %5.1f%% %5d  %s" % (t, 100 * t / tot, n, k[:110]))
