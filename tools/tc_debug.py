"""Diagnostic for the tcgen05 conv path (run on the GPU box): structured inputs that localise layout / descriptor bugs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from cfun_b200 import ops


def run(name, x, w, b=None, pad=1, relu=False, algo=ops.ALGO_TC):
    ref = F.conv3d(x, w, b, padding=pad)
    if relu:
        ref = F.relu(ref)
    ops.set_conv_algo(algo)
    try:
        y = ops.conv3d(x.cuda(), w.cuda(), b.cuda() if b is not None else None, 1, pad, relu=relu)
        torch.cuda.synchronize()
    except Exception as e:
        print("%-28s EXCEPTION %r" % (name, e))
        return
    finally:
        ops.set_conv_algo(ops.ALGO_AUTO)
    y = y.cpu()
    err = (y - ref).abs()
    rel = float(err.max() / ref.abs().max().clamp_min(1e-30))
    print("%-28s rel_err %.3e  max|ref| %.3e  nan %d" % (name, rel, float(ref.abs().max()), int(torch.isnan(y).sum())))
    if rel > 1e-4:
        bad = torch.nonzero(err > 1e-3 * ref.abs().max())
        print("   bad count", bad.shape[0], "of", err.numel(), "first:", bad[:6].tolist())
        ch = (err > 1e-3 * ref.abs().max()).sum(dim=(0, 2, 3, 4))
        print("   bad per out-channel (first 32):", ch[:32].tolist())
        sp = (err > 1e-3 * ref.abs().max()).sum(dim=(0, 1))
        print("   bad per d:", sp.sum(dim=(1, 2)).tolist())
        print("   bad per w:", sp.sum(dim=(0, 1)).tolist())
        i = bad[0].tolist()
        print("   sample got/ref:", float(y[tuple(i)]), float(ref[tuple(i)]))


def main():
    torch.manual_seed(0)
    # 1) identity: centre-tap delta kernel, y must equal x (tests A layout, swizzle, epilogue mapping)
    C = 16
    x = torch.randn(1, C, 8, 8, 16)
    w = torch.zeros(C, C, 3, 3, 3)
    for c in range(C):
        w[c, c, 1, 1, 1] = 1.0
    run("identity C16", x, w)
    # 2) shift taps: each tap alone
    for tap in [(0, 1, 1), (1, 0, 1), (1, 1, 0), (2, 2, 2)]:
        w = torch.zeros(C, C, 3, 3, 3)
        for c in range(C):
            w[c, c, tap[0], tap[1], tap[2]] = 1.0
        run("shift tap %s" % (tap,), x, w)
    # 3) channel mixing, single tap
    w = torch.zeros(C, C, 3, 3, 3)
    w[:, :, 1, 1, 1] = torch.randn(C, C)
    run("channel mix centre tap", x, w)
    # 4) full random small
    run("random C16->16", x, torch.randn(C, C, 3, 3, 3) * 0.1)
    x2 = torch.randn(1, 32, 8, 8, 16)
    run("random C32->48", x2, torch.randn(48, 32, 3, 3, 3) * 0.1)
    x3 = torch.randn(2, 20, 12, 12, 12)
    run("random C20->20 N2", x3, torch.randn(20, 20, 3, 3, 3) * 0.1)
    x4 = torch.randn(1, 128, 16, 16, 16)
    run("random 128->256 bias relu", x4, torch.randn(256, 128, 3, 3, 3) * 0.02, torch.randn(256), relu=True)
    run("single pass 128->256", x4, torch.randn(256, 128, 3, 3, 3) * 0.02, algo=ops.ALGO_TC1)
    # weight gradient on tensor cores vs torch
    for (N, Ci, S, Co, k, pd) in [(2, 20, 16, 20, 3, 1), (1, 40, 16, 40, 3, 1), (1, 128, 16, 256, 3, 1), (1, 32, 16, 24, 3, 0), (1, 80, 24, 40, 3, 1), (1, 16, 32, 16, 5, 2)]:
        x = torch.randn(N, Ci, S, S, S)
        w = (torch.randn(Co, Ci, k, k, k) * 0.05).requires_grad_(True)
        b = torch.randn(Co).requires_grad_(True)
        y = F.conv3d(x, w, b, padding=pd)
        dy = torch.randn(y.shape)
        y.backward(dy)
        for algo, nm in ((ops.ALGO_TC, "tc"), (ops.ALGO_SIMT, "simt")):
            ops.set_conv_algo(algo)
            try:
                wc = w.detach().cuda().requires_grad_(True)
                bc = b.detach().cuda().requires_grad_(True)
                yc = ops.conv3d(x.cuda(), wc, bc, 1, pd)
                yc.backward(dy.cuda())
                torch.cuda.synchronize()
                e = float((wc.grad.cpu() - w.grad).abs().max() / w.grad.abs().max())
                eb = float((bc.grad.cpu() - b.grad).abs().max() / b.grad.abs().max())
                print("wgrad %-5s N%d %d->%d @%d^3 k%d: rel_err dw %.3e db %.3e" % (nm, N, Ci, Co, S, k, e, eb))
            except Exception as ex:
                print("wgrad", nm, (N, Ci, S, Co, k), "EXC", ex)
            finally:
                ops.set_conv_algo(ops.ALGO_AUTO)
    # timing
    import time
    for (N, Ci, S, Co) in [(4, 40, 96, 40), (1, 128, 32, 256), (4, 20, 96, 20), (4, 80, 48, 80)]:
        x = ops.to_cl(torch.randn(N, Ci, S, S, S, device="cuda"))
        w = torch.randn(Co, Ci, 3, 3, 3, device="cuda") * 0.05
        for algo, nm in ((ops.ALGO_TC, "tc x3"), (ops.ALGO_TC1, "tc x1"), (ops.ALGO_SIMT, "simt")):
            ops.set_conv_algo(algo)
            try:
                ops.conv3d(x, w, None, 1, 1); torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(3):
                    ops.conv3d(x, w, None, 1, 1)
                e.record(); torch.cuda.synchronize()
                ms = s.elapsed_time(e) / 3
                fl = 2.0 * N * S ** 3 * Ci * Co * 27
                print("time %-6s N%d %d->%d @%d^3: fwd %.3f ms  %.1f TFLOP/s (incl. pack)" % (nm, N, Ci, Co, S, ms, fl / ms / 1e9))
                xg = x.clone().requires_grad_(True); wg = w.clone().requires_grad_(True)
                yy = ops.conv3d(xg, wg, None, 1, 1); gy = torch.randn_like(yy); torch.cuda.synchronize()
                s.record()
                yy.backward(gy)
                e.record(); torch.cuda.synchronize()
                print("     %-6s bwd (dgrad+wgrad) %.3f ms  debug_status %s" % (nm, s.elapsed_time(e), ops.tc_debug_status()))
                if algo == ops.ALGO_TC:
                    ref_w = wg.grad.clone()
                    ops.set_conv_algo(ops.ALGO_SIMT)
                    xg2 = x.clone().requires_grad_(True); wg2 = w.clone().requires_grad_(True)
                    ops.conv3d(xg2, wg2, None, 1, 1).backward(gy)
                    torch.cuda.synchronize()
                    print("     full-size check vs CUDA-core path: dw rel %.2e  dx rel %.2e" % (
                        float((ref_w - wg2.grad).abs().max() / wg2.grad.abs().max()),
                        float((xg.grad - xg2.grad).abs().max() / xg2.grad.abs().max())))
            except Exception as ex:
                print("time", nm, "EXC", ex)
            finally:
                ops.set_conv_algo(ops.ALGO_AUTO)


if __name__ == "__main__":
    main()
