"""Which parameter gradients of the golden 64^3 train step sit furthest from the float64 value (GPU box only).

  python tests/diag/step64_diag.py [beginning|finetune]
Prints, for the CUDA path under the current environment (CFUN_* switches), the relative deviation of every gradient norm
from the float64 norm (tests/golden/step64_fp64.npz) next to the reference's own fp32 deviation.
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from detweights import det_state
from synth import golden_step_inputs
from cfun_b200 import config as Cf, model as M

stage = sys.argv[1] if len(sys.argv) > 1 else "beginning"
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "step64_%s.npz" % stage), allow_pickle=True))
y = dict(np.load(os.path.join(ROOT, "tests", "golden", "step64_fp64.npz"), allow_pickle=True))
cfg = Cf.heart_config(64, stage, mask_pool=32, anchor_scales=(16, 32))
net = M.MaskRCNN(cfg, "/tmp/_cfun_test")
net.load_state_dict(det_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=int(g["seed_weights"])), strict=True)
net = net.cuda()
inp = golden_step_inputs(g)
net.mask.modified_u_net.injected_drop = inp["drop"]
torch.manual_seed(int(g["seed_perm"]))
dev = torch.device("cuda")
loss, losses = net.forward_backward(
    inp["image"].to(dev), None, inp["rpn_match"].to(dev)[None, :, None], inp["rpn_bbox"].to(dev)[None],
    torch.arange(1, 8).int().to(dev)[None], inp["gt_boxes"].to(dev)[None], inp["gt_masks"].to(dev)[None])
names = [str(k) for k in g["grad_names"]]
params = dict(net.named_parameters())
norms = np.array([float(params[k].grad.norm()) if params[k].grad is not None else 0.0 for k in names])
n64 = y[stage + "/grad_norms64"]
big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
dev_ours = np.abs(norms / np.maximum(n64, 1e-30) - 1)
dev_ref = np.abs(g["grad_norms"] / np.maximum(n64, 1e-30) - 1)
order = np.argsort(-np.where(big, dev_ours, 0))
print("DIAG losses ours", [float(l) for l in losses])
print("DIAG losses fp64", y[stage + "/losses"].tolist())
for i in order[:14]:
    print("DIAG %-55s ours %.2e  ref32 %.2e  norm64 %.4g" % (names[i], dev_ours[i], dev_ref[i], n64[i]))
