"""Is the tensor-pipe time data dependent (power management)?  Same conv launch on random, tiny and zero operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfun_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def med(fn, iters=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for (N, Ci, S, Co) in [(4, 40, 96, 40), (4, 20, 96, 20), (1, 128, 32, 256)]:
    for kind in ("randn", "zeros", "ones", "sparse"):
        torch.manual_seed(0)
        if kind == "randn":
            x = torch.randn(N, Ci, S, S, S, device=dev); w = torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.05
        elif kind == "zeros":
            x = torch.zeros(N, Ci, S, S, S, device=dev); w = torch.zeros(Co, Ci, 3, 3, 3, device=dev)
        elif kind == "ones":
            x = torch.ones(N, Ci, S, S, S, device=dev); w = torch.ones(Co, Ci, 3, 3, 3, device=dev)
        else:
            x = torch.randn(N, Ci, S, S, S, device=dev) * (torch.rand(N, Ci, S, S, S, device=dev) < 0.1); w = torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.05
        x = ops.to_cl(x)
        with torch.no_grad():
            t = med(lambda: ops.conv3d(x, w, None, 1, 1))
        print("N%d %d->%d @%d %-6s fwd %.3f ms" % (N, Ci, Co, S, kind, t), flush=True)
