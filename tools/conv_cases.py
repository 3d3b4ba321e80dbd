"""Time + check conv3d fwd / dgrad / wgrad of the library on a list of shapes (GPU box only).

  python tools/conv_cases.py [unet|small|all]  ->  one line per (shape, pass): ms (median of 5, L2 flushed), max error
relative to the tensor's max |value| against torch's own CUDA fp32 convolution (TF32 off), and torch's time beside it.
Each shape runs in this process; a pipeline time-out is reported through ops.tc_debug_status().
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cfun_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

UNET = [  # N, Cin, S, Cout, k, stride, pad
    (4, 20, 96, 20, 3, 1, 1), (4, 40, 96, 40, 3, 1, 1), (4, 40, 96, 20, 3, 1, 1), (4, 80, 48, 80, 3, 1, 1),
    (4, 40, 48, 40, 3, 1, 1), (4, 80, 48, 40, 3, 1, 1), (4, 160, 24, 160, 3, 1, 1), (4, 80, 24, 80, 3, 1, 1),
    (4, 20, 96, 40, 3, 2, 1), (4, 40, 48, 80, 3, 2, 1), (4, 320, 6, 320, 3, 1, 1), (4, 320, 12, 320, 3, 1, 1),
    (1, 128, 32, 256, 3, 1, 1), (1, 128, 32, 128, 3, 1, 1),
]
SMALL = [(1, 16, 16, 16, 3, 1, 1), (2, 24, 20, 40, 3, 1, 1), (1, 40, 24, 20, 3, 1, 1), (1, 80, 16, 80, 3, 1, 1),
         (2, 16, 12, 8, 3, 1, 1), (1, 48, 17, 44, 3, 1, 1), (1, 20, 32, 40, 3, 2, 1), (2, 40, 16, 80, 3, 2, 1),
         (1, 16, 48, 24, 3, 2, 1)]
TINY = [(4, 320, 6, 320, 3, 1, 1), (4, 320, 12, 320, 3, 1, 1), (4, 160, 12, 160, 3, 1, 1), (4, 320, 12, 160, 3, 1, 1),
        (4, 160, 12, 320, 3, 2, 1), (2, 32, 6, 48, 3, 1, 1), (3, 16, 5, 16, 3, 1, 1), (1, 64, 12, 32, 3, 1, 1)]
HX = [(1, 16, 16, 16, 3, 1, 1), (2, 24, 20, 40, 3, 1, 1), (1, 80, 16, 80, 3, 1, 1), (1, 48, 17, 44, 3, 1, 1), (1, 32, 16, 160, 3, 1, 1),
      (1, 128, 16, 256, 3, 1, 1), (1, 320, 8, 320, 3, 1, 1), (3, 20, 9, 20, 3, 1, 1)]
C1 = [(1, 1, 256, 16, (3, 7, 7), 2, (1, 3, 3)), (4, 1, 96, 20, 3, 1, 1), (1, 1, 40, 16, (3, 7, 7), 2, (1, 3, 3)), (2, 1, 21, 20, 3, 1, 1)]
WIDE = [(4, 160, 24, 160, 3, 1, 1), (4, 160, 24, 80, 3, 1, 1), (4, 320, 12, 320, 3, 1, 1), (4, 320, 12, 160, 3, 1, 1), (1, 128, 32, 256, 3, 1, 1),
        (1, 176, 9, 24, 3, 1, 1), (1, 128, 32, 128, 3, 1, 1), (1, 96, 16, 48, 3, 1, 1)]
PW = [(4, 40, 96, 8, 1, 1, 0), (4, 80, 48, 8, 1, 1, 0), (4, 160, 24, 8, 1, 1, 0), (1, 256, 32, 2, 1, 1, 0), (1, 256, 32, 6, 1, 1, 0),
      (2, 12, 9, 5, 1, 1, 0)]
STRIDED = [(4, 20, 96, 40, 3, 2, 1), (4, 40, 48, 80, 3, 2, 1), (4, 80, 24, 160, 3, 2, 1), (4, 160, 12, 320, 3, 2, 1)]


def med(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "unet"
    passes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fwd", "dgrad", "wgrad"]
    cases = {"unet": UNET, "small": SMALL, "all": SMALL + UNET, "strided": STRIDED, "pw": PW, "wide": WIDE, "c1": C1, "hx": HX, "tiny": TINY}[which]
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (N, Ci, S, Co, k, st, pd) in cases:
        torch.manual_seed(0)
        x = ops.to_cl(torch.randn(N, Ci, S, S, S, device=dev))
        kk = (k, k, k) if isinstance(k, int) else k
        w = torch.randn(Co, Ci, *kk, device=dev) * 0.05
        pd3 = [pd] * 3 if isinstance(pd, int) else list(pd)
        with torch.no_grad():
            yr = F.conv3d(x, w, None, st, pd)
        dy = ops.to_cl(torch.randn_like(yr))
        xr = x.detach().clone().requires_grad_(True)
        wr = w.detach().clone().requires_grad_(True)
        F.conv3d(xr, wr, None, st, pd).backward(dy)
        tag = "N%d %d->%d @%d k%s s%d" % (N, Ci, Co, S, "".join(str(v) for v in kk), st)
        if Ci == 1:
            passes = [q for q in passes if q != "dgrad"]
        for ps in passes:
            try:
                if ps == "fwd":
                    with torch.no_grad():
                        fn = lambda: ops.conv3d(x, w, None, st, pd)
                        out, ref = fn(), yr
                        tref = med(lambda: F.conv3d(x, w, None, st, pd), flush)
                elif ps == "dgrad":
                    xg = x.detach().clone().requires_grad_(True)
                    y = ops.conv3d(xg, w, None, st, pd)
                    fn = lambda: torch.autograd.grad(y, xg, dy, retain_graph=True)[0]
                    out, ref = fn(), xr.grad
                    tref = med(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [st] * 3, pd3, [1] * 3, False, [0] * 3, 1, [True, False, False]), flush)
                else:
                    wg = w.detach().clone().requires_grad_(True)
                    y = ops.conv3d(x, wg, None, st, pd)
                    fn = lambda: torch.autograd.grad(y, wg, dy, retain_graph=True)[0]
                    out, ref = fn(), wr.grad
                    tref = med(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [st] * 3, pd3, [1] * 3, False, [0] * 3, 1, [False, True, False]), flush)
                torch.cuda.synchronize()
                err = float((out - ref).abs().max() / ref.abs().max())
                t = med(fn, flush)
                if os.environ.get("CASE_KERNELS"):      # per-kernel device times of one call (CUPTI through torch.profiler)
                    from torch.profiler import profile, ProfilerActivity
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        fn()
                        torch.cuda.synchronize()
                    for ev in prof.key_averages():
                        if ev.device_time_total > 0:
                            print("    KERN %9.3f ms x%d %s" % (ev.device_time_total / 1e3, ev.count, ev.key[:90]), flush=True)
                fl = 2.0 * N * yr.shape[2] * yr.shape[3] * yr.shape[4] * Ci * Co * kk[0] * kk[1] * kk[2]
                print("CASE %-28s %-5s %8.3f ms %7.1f TF/s  err %.2e  torch %8.3f ms  dbg %s" % (
                    tag, ps, t, fl / t / 1e9, err, tref, ops.tc_debug_status()), flush=True)
            except Exception as ex:
                print("CASE %-28s %-5s EXC %s" % (tag, ps, str(ex).split("\n")[0]), flush=True)
                return


if __name__ == "__main__":
    main()
