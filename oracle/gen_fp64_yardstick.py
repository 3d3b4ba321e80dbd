"""TEST INFRASTRUCTURE: float64 run of the oracle on the golden 64^3 step -> tests/golden/step64_fp64.npz.

The deepest U-Net weight gradients are ill-conditioned in fp32 (the reference's own torch-CPU fp32 result differs from
the float64 result by ~2.5e-3 relative on mask.modified_u_net.conv_norm_lrelu_l4.0.weight), so the GPU parity test
bounds the CUDA path's distance to the float64 value by the reference's own distance instead of a flat 1e-4."""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cfun_oracle as O  # noqa: E402
from detweights import det_state  # noqa: E402
from shapes import maskrcnn_shapes  # noqa: E402
from synth import golden_step_inputs  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
out = {}
for stage in ("beginning", "finetune"):
    g = dict(np.load(os.path.join(GOLD, "step64_%s.npz" % stage)))
    sd = det_state(maskrcnn_shapes(), seed=int(g["seed_weights"]))
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    cfg = O.Cfg(image_dim=64, stage=stage, mask_pool=32, anchor_scales=(16, 32))
    inp = golden_step_inputs(g)
    keys = ["mask.modified_u_net.conv_norm_lrelu_l4.0.weight", "rpn.conv_shared.weight"]
    names = [str(k) for k in g["grad_names"]]
    leaves = {k: sd[k].clone().requires_grad_(True) for k in set(keys) | set(names)}
    sd2 = dict(sd)
    sd2.update(leaves)
    torch.manual_seed(int(g["seed_perm"]))
    res = O.train_forward(sd2, cfg, inp["image"].double(), inp["rpn_match"], inp["rpn_bbox"].double(), torch.arange(1, 8).int(),
                          inp["gt_boxes"].double(), inp["gt_masks"].double(), drop=[d.double() for d in inp["drop"]])
    res["loss"].sum().backward()
    out[stage + "/g_unet_l4"] = leaves[keys[0]].grad.flatten()[::7].numpy()
    out[stage + "/g_rpn_shared"] = leaves[keys[1]].grad.flatten()[::811].numpy()
    out[stage + "/losses"] = np.array([float(l) for l in res["losses"]])
    # float64 norm of every parameter gradient, in the order of the fp32 golden's grad_names: the yardstick for the
    # whole-step gradient check (how far is the reference's own fp32 arithmetic from the exact value, per tensor)
    out[stage + "/grad_norms64"] = np.array([float(leaves[k].grad.norm()) if leaves[k].grad is not None else 0.0 for k in names])
    big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
    print(stage, "reference fp32 vs fp64 grad norms, worst relative:", np.abs(g["grad_norms"][big] / out[stage + "/grad_norms64"][big] - 1).max())
    d = np.abs(g["g_unet_l4"] - out[stage + "/g_unet_l4"]).max() / np.abs(out[stage + "/g_unet_l4"]).max()
    print(stage, "reference fp32 vs fp64 on g_unet_l4:", d)

# U-Net layer-level gradients of the reduced-width golden model (tests/golden/layers_*.npz), float64
names = ["conv3d_c1_1", "conv3d_c3", "norm_lrelu_conv_c4.2", "conv_norm_lrelu_l4.0", "ds2_1x1_conv3d", "out_upscale_conv.1"]
keys = ["g_unet_c1_1", "g_unet_c3", "g_unet_nlc4", "g_unet_l4", "g_unet_ds2", "g_unet_up"]
for stage in ("beginning", "finetune"):
    g = dict(np.load(os.path.join(GOLD, "layers_%s.npz" % stage)))
    sd = det_state(maskrcnn_shapes(fpn=32, rpn=48, unet=4, fc=16, pool=4, num_classes=8), seed=100)
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    pre = "mask.modified_u_net."
    leaves = {pre + n + ".weight": sd[pre + n + ".weight"].clone().requires_grad_(True) for n in names}
    sd2 = dict(sd)
    sd2.update(leaves)
    drop = [torch.from_numpy(g["drop%d" % i]).double() for i in range(5)]
    y = O.unet_forward(sd2, torch.from_numpy(g["crops"]).double(), stage, drop)
    w = torch.cos(torch.arange(y.numel(), dtype=torch.float32) * 0.37).view(y.shape).double()
    (y * w).sum().backward()
    for n, k in zip(names, keys):
        gr = leaves[pre + n + ".weight"].grad
        out["layers_%s/%s" % (stage, k)] = gr.numpy() if gr is not None else np.zeros_like(g[k], dtype=np.float64)
        den = np.abs(out["layers_%s/%s" % (stage, k)]).max()
        if den > 0:
            print(stage, k, "reference fp32 vs fp64:", np.abs(g[k] - out["layers_%s/%s" % (stage, k)]).max() / den)
np.savez_compressed(os.path.join(GOLD, "step64_fp64.npz"), **out)
