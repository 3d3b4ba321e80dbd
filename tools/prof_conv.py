"""Small driver for ncu captures of the dominant conv kernels (forward / dgrad / wgrad on tensor cores)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfun_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "unet"
if which == "unet":
    N, Ci, S, Co = 4, 40, 96, 40
else:
    N, Ci, S, Co = 1, 128, 32, 256
x = ops.to_cl(torch.randn(N, Ci, S, S, S, device="cuda")).requires_grad_(True)
w = (torch.randn(Co, Ci, 3, 3, 3, device="cuda") * 0.05).requires_grad_(True)
for _ in range(2):
    y = ops.conv3d(x, w, None, 1, 1)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done", which)
