"""Where does the tensor-core forward deviate from the CUDA-core forward inside the reduced-width U-Net? (GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from detweights import det_state
from cfun_b200 import model as M, config as Cf, ops
from cfun_b200.layers import Conv3d

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "layers_beginning.npz")))
cfg = Cf.heart_config(32, "beginning", mask_pool=32, anchor_scales=(8, 16), TOP_DOWN_PYRAMID_SIZE=32, RPN_CONV_CHANNELS=48,
                      UNET_MASK_BRANCH_CHANNEL=4, FPN_CLASSIFY_FC_LAYERS_SIZE=16, POOL_SIZE=[4, 4, 4])
net = M.MaskRCNN(cfg, "/tmp/x")
net.load_state_dict(det_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=100))
net = net.cuda().train()
unet = net.mask.modified_u_net
unet.injected_drop = [torch.from_numpy(g["drop%d" % i]) for i in range(5)]
crops = torch.from_numpy(g["crops"]).cuda()
outs = {}
def hook(name):
    def f(mod, inp, out):
        outs.setdefault(name, []).append((inp[0].detach().clone(), out.detach().clone()))
    return f
for n, m in unet.named_modules():
    if isinstance(m, Conv3d):
        m.register_forward_hook(hook(n))
for mode in ("0", "1"):
    os.environ["CFUN_TC_PASSES"] = mode
    unet(crops)
torch.cuda.synchronize()
for n, lst in outs.items():
    # calls alternate: first all calls of mode 0 then mode 1 (a module applied twice has 2 calls per mode)
    k = len(lst) // 2
    for i in range(k):
        (xi0, yo0), (xi1, yo1) = lst[i], lst[k + i]
        dx = float((xi0 - xi1).abs().max() / xi0.abs().max().clamp_min(1e-30))
        dy = float((yo0 - yo1).abs().max() / yo0.abs().max().clamp_min(1e-30))
        print("%-45s call %d in %s  d_in %.2e  d_out %.2e" % (n, i, tuple(xi0.shape), dx, dy))
