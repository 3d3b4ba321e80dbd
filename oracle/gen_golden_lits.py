"""TEST INFRASTRUCTURE: golden vectors from the UNMODIFIED LiTS_2017 copy of the reference (BASELINE config 3, SURVEY.md 8f
rank 4): P3D35 backbone with the 5x7x7 stem and 24/48 planes -> FPN (160) -> RPN (320), the base-32 Modified3DUNet without
Dropout3d on non-cubic crops, the weighted mask cross-entropy [1,1,100] and the raw-Sobel edge loss.

    python oracle/gen_golden_lits.py        ->  tests/golden/lits_layers.npz, tests/golden/lits_losses.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim                      # noqa: E402
from detweights import det_state    # noqa: E402
from gen_golden import save         # noqa: E402


def lits_reference_config(R, stage, image_min, image_max, mask_pool):
    base = R["config"].Config

    class GoldenLiTS(base):
        NAME = "golden-lits"
        IMAGES_PER_GPU = 1
        NUM_CLASSES = 3
        BACKBONE = "P3D35"
        BACKBONE_STRIDES = [8, 16]
        BACKBONE_CHANNELS = [24, 48]
        FPN_CLASSIFY_FC_LAYERS_SIZE = 320
        UNET_MASK_BRANCH_CHANNEL = 32
        TOP_DOWN_PYRAMID_SIZE = 160
        RPN_CONV_CHANNELS = 320
        RPN_ANCHOR_SCALES = (16, 32)
        RPN_ANCHOR_STRIDE = 1
        RPN_ANCHOR_RATIOS = [1]
        RPN_TRAIN_ANCHORS_PER_IMAGE = 128
        PRE_NMS_LIMIT = 1000
        POST_NMS_ROIS_TRAINING = 500
        POST_NMS_ROIS_INFERENCE = 50
        USE_MINI_MASK = False
        IMAGE_RESIZE_MODE = "self"
        IMAGE_MIN_DIM = image_min
        IMAGE_MAX_DIM = image_max
        POOL_SIZE = [12, 12, 12]
        MASK_POOL_SIZE = list(mask_pool)
        DETECTION_MIN_CONFIDENCE = 0.7
        DETECTION_NMS_THRESHOLD = 0.7
        MAX_GT_INSTANCES = 32
        DETECTION_MAX_INSTANCES = 32
        TRAIN_BN = False
    return GoldenLiTS(stage)


def main():
    torch.set_num_threads(8)
    R = refshim.load("LiTS_2017")
    M = R["model"]
    cfg = lits_reference_config(R, "together", 32, 48, (32, 48, 32))
    torch.manual_seed(5)
    with refshim.quiet():
        net = M.MaskRCNN(cfg, "/tmp/_cfun_golden_lits", test_flag=False)
    for p in net.parameters():          # the LiTS model freezes the detector outside 'beginning'; goldens need its gradients
        p.requires_grad = True
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items() if k != "anchors"}
    sd = det_state(shapes, seed=300)
    missing = net.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(17)
    x = torch.randn(1, 1, 32, 48, 48, generator=g).requires_grad_(True)          # [N,C,D,H,W]: IMAGE_SHAPE (48, 48, 32)
    p2, p3 = net.fpn(x)
    lv = net.rpn(p2)
    (p2.square().sum() + p3.sum()).backward()
    res = dict(x=x.detach(), p2=p2.detach(), p3=p3.detach(), rpn_logits=lv[0].detach(), rpn_probs=lv[1].detach(), rpn_bbox=lv[2].detach(),
               g_stem=net.fpn.C1[0].weight.grad.clone(), g_P2_conv2=net.fpn.P2_conv2.weight.grad.flatten()[::11].clone(),
               g_C3_4_conv2=net.fpn.C3[4].conv2.weight.grad.clone(), gx=x.grad.clone())
    unet = net.mask.modified_u_net
    unet.train()
    crops = torch.randn(2, 1, 32, 48, 32, generator=g)
    net.zero_grad()
    y = unet(crops)
    w = torch.cos(torch.arange(y.numel(), dtype=torch.float32) * 0.37).view(y.shape)
    (y * w).sum().backward()
    res.update(crops=crops, unet_out=y.detach().flatten()[::7].clone(), unet_shape=np.array(y.shape),
               g_unet_c1_1=unet.conv3d_c1_1.weight.grad.clone(), g_unet_l4=unet.conv_norm_lrelu_l4[0].weight.grad.flatten()[::5].clone(),
               g_unet_c4=unet.conv3d_c4.weight.grad.flatten()[::13].clone())
    res["state_keys"] = np.array(sorted(shapes))
    save("lits_layers", **res)
    print("state_dict entries", len(shapes), "unet out", tuple(y.shape), "missing", missing)

    # ---- LiTS losses: weighted CE [1,1,100] (LiTS_2017/model.py:905-933) and raw-Sobel edge loss (:936-981) -------------
    P, dims = 2, (10, 12, 14)
    lab = torch.randint(0, 3, (P,) + dims, generator=g)
    tmask = torch.stack([(lab == c) for c in range(3)], 1).double()
    tcls = torch.tensor([1, 2, 0, 0])
    mlog = torch.randn(P, 3, *dims, generator=g).requires_grad_(True)
    mprob = torch.softmax(mlog, 1)
    l_mask = M.compute_mrcnn_mask_loss(tmask, tcls, mlog)
    l_edge = M.compute_mrcnn_mask_edge_loss(tmask, tcls, mprob)
    (g_edge,) = torch.autograd.grad(l_edge.sum(), mprob, retain_graph=True)
    (g_mask,) = torch.autograd.grad(l_mask, mlog)
    save("lits_losses", target_label=lab, target_class_ids=tcls, mask_logits=mlog.detach(), mask_loss=l_mask.detach(),
         edge_loss=l_edge.detach(), g_edge=g_edge, g_mask=g_mask)
    print("losses", float(l_mask), float(l_edge.sum()))


if __name__ == "__main__":
    main()
