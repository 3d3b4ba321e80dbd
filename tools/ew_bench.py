"""HBM-bound passes around the convolutions, timed one C-ABI call at a time at the U-Net / backbone shapes of the
256^3 step.  Prints achieved GB/s (algorithmic bytes: every operand read or written once) next to the measured
copy bandwidth in MEASURED_PEAKS.json.  Usage: python tools/ew_bench.py [--reps 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfun_b200 import ops  # noqa: E402
from cfun_b200.ops import _run, _ptr, _stream, empty_cl  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    peak = None
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda:0")
    shapes = [(4, 40, 96, 96, 96), (4, 20, 96, 96, 96), (4, 80, 48, 48, 48), (1, 16, 128, 128, 128), (4, 8, 96, 96, 96)]
    print(f"# copy peak {peak} GB/s")
    for (N, C, D, H, W) in shapes:
        S = D * H * W
        x = empty_cl(N, C, D, H, W, dev).normal_()
        r = empty_cl(N, C, D, H, W, dev).normal_()
        dy = empty_cl(N, C, D, H, W, dev).normal_()
        y = empty_cl(N, C, D, H, W, dev)
        dx = empty_cl(N, C, D, H, W, dev)
        dr = empty_cl(N, C, D, H, W, dev)
        a = torch.rand(N, C, device=dev) + 0.5
        b = torch.randn(N, C, device=dev)
        acc = torch.empty(2 * N * C, dtype=torch.float64, device=dev)
        mean = torch.empty(N, C, device=dev)
        rstd = torch.empty(N, C, device=dev)
        G = (C + 7) // 8
        hi = torch.empty((G, N * (D + 2), H, W, 8), dtype=torch.bfloat16, device=dev)
        lo = torch.empty_like(hi)
        tb = x.numel() * 4 / 1e9  # one fp32 tensor, GB
        st = _stream()
        cases = [
            ("instnorm_stats        (1 tensor)", 1, lambda: _run("cfun_instnorm_stats", _ptr(x), N, S, C, 1e-5, _ptr(acc),
                                                                  _ptr(mean), _ptr(rstd), st)),
            ("affine_act_fwd        (2)", 2, lambda: _run("cfun_affine_act_fwd", _ptr(x), _ptr(a), _ptr(b), C, None, _ptr(y),
                                                          N, D, H, W, C, C, 0, 1, 0.01, st)),
            ("affine_act_fwd +res   (3)", 3, lambda: _run("cfun_affine_act_fwd", _ptr(x), _ptr(a), _ptr(b), C, _ptr(r),
                                                          _ptr(y), N, D, H, W, C, C, 0, 1, 0.0, st)),
            ("affine_act_bwd +stats (3)", 3, lambda: _run("cfun_affine_act_bwd", _ptr(x), _ptr(a), _ptr(b), C, None,
                                                          _ptr(dy), _ptr(dx), None, _ptr(acc), N, D, H, W, C, C, 0, 1, 0.01,
                                                          st)),
            ("affine_act_bwd +res   (5)", 5, lambda: _run("cfun_affine_act_bwd", _ptr(x), _ptr(a), _ptr(b), C, _ptr(r),
                                                          _ptr(dy), _ptr(dx), _ptr(dr), None, N, D, H, W, C, C, 0, 1, 0.0,
                                                          st)),
            ("instnorm_bwd_apply    (3)", 3, lambda: _run("cfun_instnorm_bwd_apply", _ptr(x), _ptr(a), _ptr(b), _ptr(acc),
                                                          _ptr(dx), N, S, C, st)),
            ("pack_act_gp hi+lo P=1 (2)", 2, lambda: _run("cfun_pack_act_gp", _ptr(x), _ptr(hi), _ptr(lo), N, D, H, W, C, G, 1,
                                                          st)),
            ("torch add             (3)", 3, lambda: torch.add(x, r, out=y)),
            ("torch copy            (2)", 2, lambda: y.copy_(x)),
        ]
        print(f"shape N={N} C={C} {D}x{H}x{W}  ({tb * 1e3:.0f} MB per tensor)")
        for name, ntens, fn in cases:
            ms = timed(fn, args.reps)
            gbs = ntens * tb / (ms * 1e-3)
            frac = f"{gbs / peak:.2f}" if peak else "-"
            print(f"  {name:34s} {ms:8.3f} ms  {gbs:8.0f} GB/s  frac {frac}")
        del x, r, dy, y, dx, dr, hi, lo
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
