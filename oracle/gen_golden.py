"""TEST INFRASTRUCTURE: golden-vector generator.

Runs the *unmodified* reference (imported read-only from /root/reference through
oracle/refshim.py) on seeded inputs and writes small fixtures to tests/golden/*.npz.
Only runnable in the build container; the fixtures are committed so the CPU and GPU
test suites can run anywhere.   python oracle/gen_golden.py
"""
import os
import sys
import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from detweights import det_state  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    clean = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        clean[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **clean)
    print("wrote %-28s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def rand_boxes(rng, n, extent, smin, smax):
    c = rng.uniform(0, extent, size=(n, 3))
    s = rng.uniform(smin, smax, size=(n, 3))
    b = np.concatenate([c - s / 2, c + s / 2], 1)
    return np.clip(b, 0, extent).astype(np.float32)


def make_config(R, image_dim, stage, num_classes=8, mask_pool=32, widths=None, anchor_scales=(16, 32), **kw):
    sys.path.insert(0, refshim.REF_ROOT)
    base = R["config"].Config

    class GoldenConfig(base):
        NAME = "golden"
        IMAGES_PER_GPU = 1
        NUM_CLASSES = num_classes
        BACKBONE = "P3D19"
        BACKBONE_STRIDES = [8, 16]
        BACKBONE_CHANNELS = [16, 32]
        FPN_CLASSIFY_FC_LAYERS_SIZE = 128
        UNET_MASK_BRANCH_CHANNEL = 20
        TOP_DOWN_PYRAMID_SIZE = 128
        RPN_CONV_CHANNELS = 256
        RPN_ANCHOR_SCALES = anchor_scales
        RPN_ANCHOR_STRIDE = 1
        RPN_ANCHOR_RATIOS = [1]
        RPN_TRAIN_ANCHORS_PER_IMAGE = 128
        PRE_NMS_LIMIT = 1000
        POST_NMS_ROIS_TRAINING = 500
        POST_NMS_ROIS_INFERENCE = 64
        USE_MINI_MASK = False
        IMAGE_RESIZE_MODE = "self"
        IMAGE_MIN_DIM = image_dim
        IMAGE_MAX_DIM = image_dim
        TRAIN_ROIS_PER_IMAGE = 15
        POOL_SIZE = [12, 12, 12]
        MASK_POOL_SIZE = [mask_pool] * 3
        DETECTION_MIN_CONFIDENCE = 0.7
        DETECTION_NMS_THRESHOLD = 0.3
        MAX_GT_INSTANCES = 32
        DETECTION_MAX_INSTANCES = 32
        LOSS_WEIGHTS = {"rpn_class_loss": 100., "rpn_bbox_loss": 50., "mrcnn_class_loss": 1.,
                        "mrcnn_bbox_loss": 20., "mrcnn_mask_loss": 1., "mrcnn_mask_edge_loss": 1.}
        TRAIN_BN = False
    for k, v in (widths or {}).items():
        setattr(GoldenConfig, k, v)
    for k, v in kw.items():
        setattr(GoldenConfig, k, v)
    cfg = GoldenConfig(stage)
    cfg.MASK_SHAPE = tuple((2 if stage == "finetune" else 1) * m for m in cfg.MASK_POOL_SIZE)
    cfg.MINI_MASK_SHAPE = cfg.MASK_SHAPE
    return cfg


class InjectedDropout(nn.Module):
    """Stands in for nn.Dropout3d(p=0.6) (mask_branch.py:19) with externally drawn channel masks so the
    oracle / CUDA path can be compared on identical draws."""

    def __init__(self, masks):
        super().__init__()
        self.masks = masks
        self.i = 0

    def forward(self, x):
        m = self.masks[self.i % len(self.masks)]
        self.i += 1
        return x * m


def synth_volume(dim, cube, seed):
    """SURVEY 8d synthetic recipe at a reduced size: int16 HU-like volume [H,W,D], centred label cube."""
    rng = np.random.default_rng(seed)
    vol = np.clip(np.round(rng.normal(0, 300, size=(dim, dim, dim))), -1024, 3071).astype(np.int16)
    lab = np.zeros((dim, dim, dim), dtype=np.int32)
    a = (dim - cube) // 2
    lab[a:a + cube, a:a + cube, a:a + cube] = rng.integers(1, 8, size=(cube, cube, cube))
    return vol, lab


def main():
    torch.set_num_threads(8)
    R = refshim.load()
    M, U = R["model"], R["utils"]
    rng = np.random.default_rng(7)

    # ---- anchors (utils.py:467-528), D != H != W to pin the y-major order -------------------------
    shapes = M.compute_backbone_shapes(type("c", (), {"BACKBONE_STRIDES": [8, 16]}), (48, 64, 32, 1))
    anc = U.generate_pyramid_anchors((16, 32), [1], shapes, [8, 16], 1)
    save("anchors", image_shape=np.array([48, 64, 32, 1]), scales=np.array([16, 32]), strides=np.array([8, 16]),
         shapes=shapes, anchors=anc)

    # ---- NMS (utils.py:122-157) -------------------------------------------------------------------
    cases = {}
    b = rand_boxes(rng, 400, 64, 6, 30); s = rng.permutation(400).astype(np.float32) / 400
    cases["rand_t7_m50"] = (b, s, 0.7, 50)
    cases["rand_t3_all"] = (b, s, 0.3, 400)
    cases["rand_t5_m1"] = (b, s, 0.5, 1)
    nb = np.array([[10, 10, 10, 30, 30, 30], [12, 12, 12, 28, 28, 28], [10, 10, 10, 30, 30, 30.5],
                   [0, 0, 0, 0, 5, 5], [40, 40, 40, 41, 41, 41], [10, 10, 10, 20, 30, 30]], dtype=np.float32)
    cases["nested_degenerate"] = (nb, np.array([.9, .8, .7, .6, .5, .4], dtype=np.float32), 0.5, 10)
    bi = np.round(rand_boxes(rng, 300, 32, 4, 16)); si = rng.permutation(300).astype(np.float32)
    cases["integer_boxes_t3"] = (bi.astype(np.float32), si, 0.3, 300)
    out = {}
    for k, (bb, ss, t, m) in cases.items():
        out[k + "/boxes"] = bb; out[k + "/scores"] = ss; out[k + "/thr"] = np.float64(t); out[k + "/max"] = np.int64(m)
        out[k + "/keep"] = U.non_max_suppression(bb, ss, t, m)
        out[k + "/iou0"] = U.compute_iou(bb[0], bb, ((bb[0, 3] - bb[0, 0]) * (bb[0, 4] - bb[0, 1]) * (bb[0, 5] - bb[0, 2])),
                                         (bb[:, 3] - bb[:, 0]) * (bb[:, 4] - bb[:, 1]) * (bb[:, 5] - bb[:, 2]))
    save("nms", **out)

    # ---- box decode / clip / proposal layer (model.py:155-258) ------------------------------------
    cfg = make_config(R, 64, "beginning")
    anchors = torch.from_numpy(U.generate_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, [1],
                               M.compute_backbone_shapes(cfg, cfg.IMAGE_SHAPE), cfg.BACKBONE_STRIDES, 1)).float()
    A = anchors.shape[0]
    g = torch.Generator().manual_seed(11)
    probs = torch.softmax(torch.randn(A, 2, generator=g) * 2, dim=1)
    deltas = torch.randn(A, 6, generator=g) * 0.5
    dec = M.apply_box_deltas(anchors.clone(), deltas * 0.1)
    clp = M.clip_boxes(dec, np.array([0, 0, 0, 64, 64, 64], dtype=np.float32))
    with refshim.quiet():
        rois_tr = M.proposal_layer([probs[None].clone(), deltas[None].clone()], 500, 0.7, anchors, cfg)
        rois_inf = M.proposal_layer([probs[None].clone(), deltas[None].clone()], 64, 0.7, anchors, cfg)
    save("proposal", anchors=anchors, probs=probs, deltas=deltas, decoded=dec, clipped=clp, rois_training=rois_tr[0],
         rois_inference=rois_inf[0], image_shape=np.array(cfg.IMAGE_SHAPE))

    # ---- RoI crop-resize + pyramid level (model.py:265-370) ---------------------------------------
    g = torch.Generator().manual_seed(12)
    f2 = torch.randn(1, 3, 16, 20, 12, generator=g)
    f3 = torch.randn(1, 3, 8, 10, 6, generator=g)
    bx = torch.rand(14, 3, generator=g) * 0.6
    sz = torch.rand(14, 3, generator=g) * 0.5 + 0.02
    boxes = torch.cat([bx, torch.clamp(bx + sz, max=1.0)], 1)
    boxes[0] = torch.tensor([0.25, 0.1, 0.1, 0.25, 0.5, 0.5])          # empty crop -> zeros row
    boxes[1] = torch.tensor([0.0, 0.0, 0.0, 1.0, 1.0, 1.0])          # whole map
    boxes[2] = torch.tensor([0.2, 0.2, 0.2, 0.2 + 0.3536, 0.2 + 0.3536, 0.2 + 0.3536])  # near level boundary
    boxes[3] = torch.tensor([0.5, 0.5, 0.5, 0.56, 0.56, 0.56])       # single-voxel crop
    with refshim.quiet():
        pooled = M.pyramid_roi_align([boxes.clone(), f2.clone(), f3.clone()], [4, 5, 3])
        single = M.RoI_Align(f2[0], [4, 5, 3], boxes.clone())
    lvl = (4 + (1. / 3.) * M.log2((boxes[:, 4] - boxes[:, 1]) * (boxes[:, 5] - boxes[:, 2]) * (boxes[:, 3] - boxes[:, 0]))).round().int().clamp(2, 3)
    save("roialign", f2=f2[0], f3=f3[0], boxes=boxes, pool=np.array([4, 5, 3]), pooled=pooled, single_level=single, level=lvl)

    # ---- overlaps / refinement / detection targets (model.py:377-563, utils.py:92-119) ------------
    g = torch.Generator().manual_seed(13)
    jit = np.array([0.3, 0.3, 0.3, 0.7, 0.7, 0.7], dtype=np.float32) + rng.uniform(-0.06, 0.06, size=(30, 6)).astype(np.float32)
    props = torch.from_numpy(np.concatenate([rand_boxes(rng, 90, 1.0, 0.2, 0.6), jit], 0)[rng.permutation(120)])
    gtb = torch.tensor([[0.3, 0.3, 0.3, 0.7, 0.7, 0.7]]).repeat(7, 1)
    ov = M.bbox_overlaps(props, gtb)
    refin = U.box_refinement(props[:20], gtb[:1].repeat(20, 1))
    lab = np.zeros((24, 24, 24), dtype=np.int32)
    lab[7:17, 7:17, 7:17] = rng.integers(1, 8, size=(10, 10, 10))
    gt_masks = torch.from_numpy(np.stack([(lab == c) for c in range(8)]).astype(np.float32))
    dcfg = make_config(R, 64, "beginning", mask_pool=16)
    torch.manual_seed(1234)
    with refshim.quiet():
        p_rois, rois, cls, dl, msk = M.detection_target_layer(props[None], torch.arange(1, 8)[None].int(), gtb[None],
                                                              gt_masks[None], dcfg)
    save("dtl", proposals=props, gt_boxes=gtb, overlaps=ov, refinement=refin, label=lab, seed=np.int64(1234),
         positive_rois=p_rois, rois=rois, class_ids=cls, deltas=dl, masks=msk.numpy().astype(np.uint8),
         mask_shape=np.array(dcfg.MASK_SHAPE))

    # ---- nearest-neighbour resize stand-in (utils.py:318-339; parity unpinned vs skimage) ---------
    a = rng.integers(0, 9, size=(3, 7, 11, 5)).astype(np.float64)
    save("nnresize", src=a, dst=refshim._skimage_resize(a, (3, 16, 16, 16), order=0), dst_small=refshim._skimage_resize(a, (3, 4, 5, 3), order=0))

    # ---- detection refinement, inference (model.py:584-676) ---------------------------------------
    g = torch.Generator().manual_seed(14)
    r_rois = torch.from_numpy(rand_boxes(rng, 64, 1.0, 0.2, 0.5))
    r_probs = torch.softmax(torch.randn(64, 2, generator=g) * 3, dim=1)
    r_delta = torch.randn(64, 2, 6, generator=g) * 0.3
    icfg = make_config(R, 64, "beginning")
    with refshim.quiet():
        det = M.refine_detections(r_rois.clone(), r_probs, r_delta, np.array([0, 0, 0, 64, 64, 64]), icfg)
    save("refine", rois=r_rois, probs=r_probs, deltas=r_delta, detections=det)

    # ---- layers: backbone+FPN, RPN, classifier, U-Net (both stages, dropout injected) -------------
    small = {"TOP_DOWN_PYRAMID_SIZE": 32, "RPN_CONV_CHANNELS": 48, "UNET_MASK_BRANCH_CHANNEL": 4,
             "FPN_CLASSIFY_FC_LAYERS_SIZE": 16, "POOL_SIZE": [4, 4, 4]}
    for stage in ("beginning", "finetune"):
        cfg = make_config(R, 32, stage, mask_pool=32, widths=small, anchor_scales=(8, 16))
        torch.manual_seed(3)
        with refshim.quiet():
            net = M.MaskRCNN(cfg, "/tmp/_cfun_golden")
        sd = det_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=100)
        net.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(21)
        x = torch.randn(1, 1, 32, 32, 32, generator=g)
        net.train()
        for m in net.modules():
            if isinstance(m, nn.BatchNorm3d):
                m.eval()
        xg = x.clone().requires_grad_(True)
        p2, p3 = net.fpn(xg)
        (p2.square().sum() + p3.sum()).backward()
        gw = net.fpn.P2_conv2.weight.grad.clone(); gstem = net.fpn.C1[0].weight.grad.clone()
        gc2 = net.fpn.C2[1].conv2.weight.grad.clone()
        logits, probs, bbox = net.rpn(p2.detach())
        unet = net.mask.modified_u_net
        crops = torch.randn(1, 1, 32, 32, 32, generator=g)
        chans = [4, 8, 16, 32, 64]
        drop = [(torch.rand(1, c, 1, 1, 1, generator=g) > 0.6).float() / 0.4 for c in chans]
        unet.dropout3d = InjectedDropout(drop)
        net.zero_grad()
        y = unet(crops)
        w = torch.cos(torch.arange(y.numel(), dtype=torch.float32) * 0.37).view(y.shape)
        (y * w).sum().backward()
        ug = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in unet.named_parameters()}
        unet.eval(); unet.dropout3d = nn.Identity()
        y_eval = unet(crops)
        pooled = torch.randn(5, 32, 4, 4, 4, generator=g)
        c_logits, c_probs, c_bbox = None, None, None
        x1 = net.classifier.relu(net.classifier.bn1(net.classifier.conv1(pooled)))
        x1 = net.classifier.relu(net.classifier.bn2(net.classifier.conv2(x1))).view(-1, 16)
        c_logits = net.classifier.linear_class(x1); c_bbox = net.classifier.linear_bbox(x1).view(5, -1, 6)
        save("layers_" + stage, x=x, p2=p2, p3=p3, g_P2_conv2=gw, g_stem=gstem, g_C2_1_conv2=gc2, gx=xg.grad,
             rpn_logits=logits, rpn_probs=probs, rpn_bbox=bbox, crops=crops,
             **{"drop%d" % i: d for i, d in enumerate(drop)}, unet_train=y.flatten()[::13], unet_eval=y_eval.flatten()[::13],
             g_unet_c1_1=ug["conv3d_c1_1.weight"], g_unet_c3=ug["conv3d_c3.weight"],
             g_unet_nlc4=ug["norm_lrelu_conv_c4.2.weight"], g_unet_l4=ug["conv_norm_lrelu_l4.0.weight"],
             g_unet_ds2=ug["ds2_1x1_conv3d.weight"], g_unet_up=ug["out_upscale_conv.1.weight"],
             pooled=pooled, cls_logits=c_logits, cls_bbox=c_bbox, seed=np.int64(100))

    # ---- losses incl. the Sobel edge loss (model.py:808-981) --------------------------------------
    g = torch.Generator().manual_seed(31)
    P, Mx = 2, 12
    tmask_lab = torch.randint(0, 8, (P, Mx, Mx, Mx), generator=g)
    tmask = torch.stack([(tmask_lab == c) for c in range(8)], 1).double()
    tcls = torch.tensor([3, 5, 0, 0, 0])
    mlog = torch.randn(P, 8, Mx, Mx, Mx, generator=g).requires_grad_(True)
    mprob = torch.softmax(mlog, 1)
    l_mask = M.compute_mrcnn_mask_loss(tmask, tcls, mlog)
    l_edge = M.compute_mrcnn_mask_edge_loss(tmask, tcls, mprob)
    (g_edge,) = torch.autograd.grad(l_edge.sum(), mprob, retain_graph=True)
    (g_mask,) = torch.autograd.grad(l_mask, mlog)
    Aa = 200
    rmatch = torch.from_numpy(rng.choice([-1, 0, 1], size=(1, Aa, 1), p=[.3, .5, .2]).astype(np.int32))
    rlog = torch.randn(1, Aa, 2, generator=g); rbb = torch.randn(1, Aa, 6, generator=g)
    npos = int((rmatch == 1).sum())
    rtgt = torch.zeros(1, 128, 6); rtgt[0, :npos] = torch.randn(npos, 6, generator=g)
    clog = torch.randn(5, 2, generator=g); cbb = torch.randn(5, 2, 6, generator=g); tdel = torch.randn(5, 6, generator=g)
    bin_ids = (tcls > 0).long()
    save("losses", target_label=tmask_lab, target_class_ids=tcls, mask_logits=mlog, mask_loss=l_mask, edge_loss=l_edge,
         g_edge=g_edge, g_mask=g_mask, rpn_match=rmatch, rpn_logits=rlog, rpn_bbox=rbb, rpn_target=rtgt,
         rpn_class_loss=M.compute_rpn_class_loss(rmatch, rlog), rpn_bbox_loss=M.compute_rpn_bbox_loss(rtgt, rmatch, rbb),
         cls_logits=clog, cls_bbox=cbb, target_deltas=tdel, cls_loss=M.compute_mrcnn_class_loss(bin_ids, clog),
         bbox_loss=M.compute_mrcnn_bbox_loss(tdel, bin_ids, cbb))

    # ---- one whole train step at 64^3, real layer widths, both stages ------------------------------
    for stage, mp in (("beginning", 32), ("finetune", 32)):
        cfg = make_config(R, 64, stage, mask_pool=mp, anchor_scales=(16, 32))
        torch.manual_seed(5)
        with refshim.quiet():
            net = M.MaskRCNN(cfg, "/tmp/_cfun_golden")
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(det_state(shapes, seed=200), strict=True)
        vol, lab = synth_volume(64, 18, seed=1000)
        image = M.mold_image(vol.astype(np.float32)[..., None]).transpose((3, 2, 0, 1))[None]    # [1,1,D,H,W]
        labt = lab.transpose((2, 0, 1))
        a = (64 - 18) // 2
        gt_box = np.array([[a, a, a, a + 18, a + 18, a + 18]], dtype=np.int32)
        gt_boxes = np.tile(gt_box, (7, 1))
        gt_masks = np.stack([(labt == c) for c in range(8)]).astype(np.float32)
        np.random.seed(9)
        rpn_match, rpn_bbox = M.build_rpn_targets(net.anchors.numpy(), gt_box, cfg)
        g = torch.Generator().manual_seed(41)
        chans = [20, 40, 80, 160, 320]
        drop = None
        res = {}
        for attempt in range(1):
            torch.manual_seed(77)
            net.zero_grad()
            # injected dropout sized for the positive count (drawn generously, sliced by the module)
            drop = [(torch.rand(4, c, 1, 1, 1, generator=g) > 0.6).float() / 0.4 for c in chans]

            class SlicedDrop(InjectedDropout):
                def forward(self, x):
                    m = self.masks[self.i % len(self.masks)][:x.shape[0]]
                    self.i += 1
                    return x * m
            net.mask.modified_u_net.dropout3d = SlicedDrop(drop)
            with refshim.quiet():
                outs = net.predict([torch.from_numpy(image).float(), None, torch.arange(1, 8)[None].int(),
                                    torch.from_numpy(gt_boxes).float()[None], torch.from_numpy(gt_masks)[None]], "training")
                rpn_class_logits, rpn_pred_bbox, tcls, c_logits, tdel, c_bbox, tmask, m_probs, m_logits = outs
                losses = M.compute_losses(torch.from_numpy(rpn_match[None, :, None]), torch.from_numpy(rpn_bbox).float()[None],
                                          rpn_class_logits, rpn_pred_bbox, tcls, c_logits, tdel, c_bbox, tmask, m_probs,
                                          m_logits, cfg.STAGE)
            w = cfg.LOSS_WEIGHTS
            total = sum(w[k] * l for k, l in zip(["rpn_class_loss", "rpn_bbox_loss", "mrcnn_class_loss", "mrcnn_bbox_loss",
                                                  "mrcnn_mask_loss", "mrcnn_mask_edge_loss"], losses))
            total.sum().backward()
            names = [k for k, p in net.named_parameters() if p.requires_grad]
            gnorms = np.array([float(p.grad.norm()) if p.grad is not None else 0.0
                               for k, p in net.named_parameters() if p.requires_grad])
            print(stage, "positives", int((tcls > 0).sum()), "rois", int(tcls.shape[0]))
            g_rpn = net.rpn.conv_shared.weight.grad.flatten()[::811].clone()
            g_l4 = net.mask.modified_u_net.conv_norm_lrelu_l4[0].weight.grad.flatten()[::7].clone()
            pre = float(torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0))
        print(stage, "positives", int((tcls > 0).sum()), "rois", int(tcls.shape[0]), "losses", [float(l) for l in losses], "gnorm", pre)
        save("step64_" + stage, vol=vol, label=lab, rpn_match=rpn_match, rpn_bbox=rpn_bbox, seed_weights=np.int64(200),
             seed_perm=np.int64(77), **{"drop%d" % i: d for i, d in enumerate(drop)},
             losses=np.array([float(l) for l in losses]), total=float(total), target_class_ids=tcls,
             target_deltas=tdel, rois_count=np.int64(tcls.shape[0]), grad_names=np.array(names), grad_norms=gnorms,
             grad_total_norm=np.float64(pre), rpn_class_logits=rpn_class_logits[0, ::37], rpn_pred_bbox=rpn_pred_bbox[0, ::37],
             mask_logits_sample=m_logits.detach().flatten()[::997] if m_logits.numel() else np.zeros(0),
             g_rpn_shared=g_rpn, g_unet_l4=g_l4)


if __name__ == "__main__":
    main()
