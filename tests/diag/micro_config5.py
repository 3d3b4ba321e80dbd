"""BASELINE.json config 5 micro-benchmark: 3-D NMS and RoI crop-resize on 10 000 random proposals over a
256^3 map, one GPU.  CUDA-event medians of 5 runs; the numpy / torch-CPU restatement timed beside them."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import cfun_oracle as O
from cfun_b200 import ops

def med(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

rng = np.random.default_rng(0)
n = 10000
c = rng.uniform(0, 256, size=(n, 3)); s = rng.uniform(16, 128, size=(n, 3))
b = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.float32)
sc = rng.uniform(0, 1, size=n).astype(np.float32)
bd, sd = torch.from_numpy(b).cuda(), torch.from_numpy(sc).cuda()
for thr, mx in ((0.7, 500), (0.7, n), (0.3, n)):
    def run():
        order = ops.sort_desc(sd)
        return ops.nms3d(bd[order.long()], thr, mx)
    t = med(run)
    keep, cnt = run()
    t0 = time.time(); ref = O.non_max_suppression(b, sc, thr, mx); tc = (time.time() - t0) * 1e3
    print("NMS n=%d thr=%.1f max=%d: kept %d  GPU sort+NMS %.3f ms  numpy %.1f ms  (x%.0f)" % (n, thr, mx, int(cnt), t, tc, tc / t))
boxes = torch.from_numpy(b / 256.0).cuda()
fmap = torch.randn(1, 1, 256, 256, 256, device="cuda")
t = med(lambda: ops.roi_crop_resize(fmap, None, boxes, None, (12, 12, 12), True))
out_mb = n * 12 ** 3 * 4 / 1e6
print("RoI crop-resize C=1 256^3, 10000 boxes, pool 12^3: %.3f ms (%.1f MB out, gather-minimal input %.0f MB -> %.0f GB/s)" % (
    t, out_mb, n * 8 * 1728 * 4 / 1e6, (out_mb + n * 8 * 1728 * 4 / 1e6) / t))
t = med(lambda: ops.roi_crop_resize(fmap, None, boxes[:16], None, (96, 96, 96), True))
print("RoI crop-resize C=1 256^3, 16 boxes, pool 96^3: %.3f ms (%.1f MB out)" % (t, 16 * 96 ** 3 * 4 / 1e6))
f2 = torch.randn(1, 128, 32, 32, 32, device="cuda")
t = med(lambda: ops.roi_crop_resize(f2, None, boxes[:1000], None, (12, 12, 12), True))
print("RoI crop-resize C=128 32^3, 1000 boxes, pool 12^3 (classifier case): %.3f ms (%.1f MB out)" % (t, 1000 * 128 * 1728 * 4 / 1e6))
t0 = time.time(); O.roi_align(fmap[0].cpu(), (12, 12, 12), boxes[:500].cpu()); tc = (time.time() - t0) * 1e3
print("CPU restatement, 500 boxes pool 12^3: %.1f ms (%.3f ms / box)" % (tc, tc / 500))
