"""one tcgen05 weight-gradient case in its own process (a failing kernel poisons the CUDA context)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cfun_b200 import ops
N, Ci, D, H, W, Co, k, pd = [int(v) for v in sys.argv[1:9]]
mode = sys.argv[9] if len(sys.argv) > 9 else "wgrad"
torch.manual_seed(0)
x = torch.randn(N, Ci, D, H, W)
w = (torch.randn(Co, Ci, k, k, k) * 0.05).requires_grad_(True)
y = F.conv3d(x, w, None, padding=pd)
dy = torch.randn(y.shape)
y.backward(dy)
ops.set_conv_algo(ops.ALGO_TC)
try:
    wc = w.detach().cuda().requires_grad_(mode == "wgrad")
    xc = x.cuda().requires_grad_(mode != "wgrad")
    yc = ops.conv3d(xc, wc, None, 1, pd)
    yc.backward(dy.cuda())
    torch.cuda.synchronize()
    print("CASE tc_debug_status:", ops.tc_debug_status())
    ef = float((yc.detach().cpu() - y.detach()).abs().max() / y.abs().max())
    if mode == "wgrad":
        e = float((wc.grad.cpu() - w.grad).abs().max() / w.grad.abs().max())
        print("CASE %s wgrad rel_err %.3e fwd %.3e" % (sys.argv[1:9], e, ef))
    else:
        print("CASE %s fwd rel_err %.3e" % (sys.argv[1:9], ef))
except Exception as ex:
    print("CASE %s EXC %s" % (sys.argv[1:9], str(ex).split("\n")[0]))
