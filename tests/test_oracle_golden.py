"""Pins the CPU oracle (oracle/cfun_oracle.py) against golden vectors emitted by the unmodified
reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import torch
import pytest

import cfun_oracle as O
from detweights import det_state
from conftest import load_golden, rel_err

T = torch.from_numpy


def test_anchor_order_matches_reference():
    g = load_golden("anchors")
    shapes = O.backbone_shapes(tuple(g["image_shape"]), tuple(g["strides"]))
    assert np.array_equal(shapes, g["shapes"])
    a = O.generate_pyramid_anchors(tuple(g["scales"]), shapes, tuple(g["strides"]), 1)
    assert a.shape == g["anchors"].shape
    assert np.array_equal(a, g["anchors"])          # exact integers / halves


@pytest.mark.parametrize("case", ["rand_t7_m50", "rand_t3_all", "rand_t5_m1", "nested_degenerate", "integer_boxes_t3"])
def test_nms_bit_exact(case):
    g = load_golden("nms")
    b, s = g[case + "/boxes"], g[case + "/scores"]
    keep = O.non_max_suppression(b, s, float(g[case + "/thr"]), int(g[case + "/max"]))
    assert keep.dtype == np.int32
    assert np.array_equal(keep, g[case + "/keep"])
    vol = O.box_volume(b)
    iou = O.compute_iou(b[0], b, vol[0], vol)
    assert np.array_equal(iou.view(np.uint32), g[case + "/iou0"].view(np.uint32))   # bit pattern


def test_decode_clip_proposals():
    g = load_golden("proposal")
    anchors, probs, deltas = T(g["anchors"]), T(g["probs"]), T(g["deltas"])
    dec = O.apply_box_deltas(anchors, deltas * 0.1)
    assert np.array_equal(dec.numpy(), g["decoded"])
    clp = O.clip_boxes(dec, [0, 0, 0, 64, 64, 64])
    assert np.array_equal(clp.numpy(), g["clipped"])
    for key, count in (("rois_training", 500), ("rois_inference", 64)):
        rois, keep, order = O.proposal_layer(probs, deltas, anchors, count, 0.7, 1000, tuple(g["image_shape"]))
        assert np.array_equal(rois.numpy(), g[key])


def test_roi_crop_resize_and_levels():
    g = load_golden("roialign")
    boxes = T(g["boxes"])
    assert np.array_equal(O.roi_level(boxes).numpy(), g["level"])
    single = O.roi_align(T(g["f2"]), tuple(g["pool"]), boxes)
    assert np.array_equal(single.numpy(), g["single_level"])
    pooled = O.pyramid_roi_align(boxes, [T(g["f2"]), T(g["f3"])], tuple(g["pool"]))
    assert np.array_equal(pooled.numpy(), g["pooled"])
    assert np.all(g["pooled"][0] == 0)      # the empty crop row stays zero (model.py:281-287)


def test_overlaps_refinement_targets():
    g = load_golden("dtl")
    props, gtb = T(g["proposals"]), T(g["gt_boxes"])
    assert np.array_equal(O.bbox_overlaps(props, gtb).numpy(), g["overlaps"])
    assert np.array_equal(O.box_refinement(props[:20], gtb[:1].repeat(20, 1)).numpy(), g["refinement"])
    lab = g["label"]
    gt_masks = T(np.stack([(lab == c) for c in range(8)]).astype(np.float32))
    iou_max = O.bbox_overlaps(props, gtb).max(1)[0]
    npos, nneg = int((iou_max >= 0.5).sum()), int((iou_max < 0.5).sum())
    torch.manual_seed(int(g["seed"]))
    perm_pos = torch.randperm(npos)
    perm_neg = torch.randperm(nneg)
    p_rois, rois, cls, dl, msk = O.detection_target_layer(
        props, torch.arange(1, 8).int(), gtb, gt_masks, tuple(g["mask_shape"]), 15, 0.33, 0.5,
        (0.1, 0.1, 0.1, 0.2, 0.2, 0.2), perm_pos, perm_neg)
    assert np.array_equal(p_rois.numpy(), g["positive_rois"])
    assert np.array_equal(rois.numpy(), g["rois"])
    assert np.array_equal(cls.numpy(), g["class_ids"])
    assert np.array_equal(dl.numpy(), g["deltas"])
    assert msk.dtype == torch.float64
    assert np.array_equal(msk.numpy().astype(np.uint8), g["masks"])


def test_nn_resize_matches_scipy_stand_in():
    g = load_golden("nnresize")
    assert np.array_equal(O.nn_resize(g["src"], (3, 16, 16, 16)), g["dst"])
    assert np.array_equal(O.nn_resize(g["src"], (3, 4, 5, 3)), g["dst_small"])


def test_refine_detections():
    g = load_golden("refine")
    det = O.refine_detections(T(g["rois"]), T(g["probs"]), T(g["deltas"]), [0, 0, 0, 64, 64, 64], (64, 64, 64, 1),
                              0.7, 0.3, 32)
    assert np.array_equal(det.numpy(), g["detections"])


SMALL_SHAPES = None


def _small_state(stage):
    """state_dict shapes of the reduced-width golden model (see gen_golden.py 'small')."""
    from shapes import maskrcnn_shapes
    return det_state(maskrcnn_shapes(fpn=32, rpn=48, unet=4, fc=16, pool=4, num_classes=8), seed=100)


@pytest.mark.parametrize("stage", ["beginning", "finetune"])
def test_layers_forward_backward(stage):
    g = load_golden("layers_" + stage)
    sd = _small_state(stage)
    x = T(g["x"]).requires_grad_(True)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in ("fpn.P2_conv2.weight", "fpn.C1.0.weight", "fpn.C2.1.conv2.weight")}
    sd2 = dict(sd); sd2.update(leaves)
    p2, p3 = O.fpn_forward(sd2, x)
    assert rel_err(p2.detach().numpy(), g["p2"]) < 1e-5 and rel_err(p3.detach().numpy(), g["p3"]) < 1e-5
    (p2.square().sum() + p3.sum()).backward()
    assert rel_err(leaves["fpn.P2_conv2.weight"].grad.numpy(), g["g_P2_conv2"]) < 1e-4
    assert rel_err(leaves["fpn.C1.0.weight"].grad.numpy(), g["g_stem"]) < 1e-4
    assert rel_err(leaves["fpn.C2.1.conv2.weight"].grad.numpy(), g["g_C2_1_conv2"]) < 1e-4
    assert rel_err(x.grad.numpy(), g["gx"]) < 1e-4
    logits, probs, bbox = O.rpn_forward(sd, p2.detach())
    assert rel_err(logits.numpy(), g["rpn_logits"]) < 1e-5
    assert rel_err(probs.numpy(), g["rpn_probs"]) < 1e-5
    assert rel_err(bbox.numpy(), g["rpn_bbox"]) < 1e-5
    # U-Net, train mode with the injected dropout draws, then eval
    pre = "mask.modified_u_net."
    names = ["conv3d_c1_1", "conv3d_c3", "norm_lrelu_conv_c4.2", "conv_norm_lrelu_l4.0", "ds2_1x1_conv3d", "out_upscale_conv.1"]
    leaves = {pre + n + ".weight": sd[pre + n + ".weight"].clone().requires_grad_(True) for n in names}
    sd2 = dict(sd); sd2.update(leaves)
    drop = [T(g["drop%d" % i]) for i in range(5)]
    y = O.unet_forward(sd2, T(g["crops"]), stage, drop)
    assert rel_err(y.detach().flatten()[::13].numpy(), g["unet_train"]) < 1e-4
    w = torch.cos(torch.arange(y.numel(), dtype=torch.float32) * 0.37).view(y.shape)
    (y * w).sum().backward()
    for n, key in zip(names, ["g_unet_c1_1", "g_unet_c3", "g_unet_nlc4", "g_unet_l4", "g_unet_ds2", "g_unet_up"]):
        gr = leaves[pre + n + ".weight"].grad
        gr = gr.numpy() if gr is not None else np.zeros_like(g[key])
        assert rel_err(gr, g[key]) < 2e-4, key
    y_eval = O.unet_forward(sd, T(g["crops"]), stage, None)
    assert rel_err(y_eval.flatten()[::13].numpy(), g["unet_eval"]) < 1e-4
    c_logits, c_probs, c_bbox = O.classifier_forward(sd, T(g["pooled"]))
    assert rel_err(c_logits.numpy(), g["cls_logits"]) < 1e-5 and rel_err(c_bbox.numpy(), g["cls_bbox"]) < 1e-5


def test_losses_and_edge_loss():
    g = load_golden("losses")
    lab = T(g["target_label"])
    tmask = torch.stack([(lab == c) for c in range(8)], 1).double()
    tcls = T(g["target_class_ids"])
    mlog = T(g["mask_logits"]).requires_grad_(True)
    mprob = torch.softmax(mlog, 1)
    l_mask = O.mrcnn_mask_loss(tmask, tcls, mlog)
    l_edge = O.mrcnn_mask_edge_loss(tmask, tcls, mprob)
    assert abs(float(l_mask) - float(g["mask_loss"])) < 1e-6 * abs(float(g["mask_loss"])) + 1e-7
    assert abs(float(l_edge) - float(g["edge_loss"])) < 1e-5 * abs(float(g["edge_loss"]))
    (g_edge,) = torch.autograd.grad(l_edge.sum(), mprob, retain_graph=True)
    (g_mask,) = torch.autograd.grad(l_mask, mlog)
    assert rel_err(g_edge.numpy(), g["g_edge"]) < 1e-5
    assert rel_err(g_mask.numpy(), g["g_mask"]) < 1e-5
    rmatch = T(g["rpn_match"])[0, :, 0]
    assert abs(float(O.rpn_class_loss(rmatch, T(g["rpn_logits"])[0])) - float(g["rpn_class_loss"])) < 1e-6
    assert abs(float(O.rpn_bbox_loss(T(g["rpn_target"])[0], rmatch, T(g["rpn_bbox"])[0])) - float(g["rpn_bbox_loss"])) < 1e-6
    bin_ids = (tcls > 0).long()
    assert abs(float(O.mrcnn_class_loss(bin_ids, T(g["cls_logits"]))) - float(g["cls_loss"])) < 1e-6
    assert abs(float(O.mrcnn_bbox_loss(T(g["target_deltas"]), bin_ids, T(g["cls_bbox"]))) - float(g["bbox_loss"])) < 1e-6


@pytest.mark.parametrize("stage", ["beginning", "finetune"])
def test_whole_step_64(stage):
    from shapes import maskrcnn_shapes
    from synth import golden_step_inputs
    g = load_golden("step64_" + stage)
    shapes = maskrcnn_shapes()
    sd = det_state(shapes, seed=int(g["seed_weights"]))
    cfg = O.Cfg(image_dim=64, stage=stage, mask_pool=32, anchor_scales=(16, 32))
    inp = golden_step_inputs(g)
    trainable = [k for k in g["grad_names"]]
    torch.manual_seed(int(g["seed_perm"]))
    out, grads, norm = O.train_step_grads(sd, cfg, trainable, inp["image"], inp["rpn_match"], inp["rpn_bbox"],
                                          torch.arange(1, 8).int(), inp["gt_boxes"], inp["gt_masks"],
                                          drop=[d[:1] for d in inp["drop"]])
    losses = np.array([float(l) for l in out["losses"]])
    assert np.allclose(losses, g["losses"], rtol=2e-4, atol=1e-6), (losses, g["losses"])
    assert np.array_equal(out["target_class_ids"].numpy(), g["target_class_ids"])
    assert abs(norm - float(g["grad_total_norm"])) < 1e-3 * float(g["grad_total_norm"])
    assert rel_err(out["rpn_class_logits"][::37].detach().numpy(), g["rpn_class_logits"]) < 1e-4


def test_c_restatement_of_nms_agrees_with_numpy_oracle_and_goldens():
    import ctypes, os, subprocess
    from conftest import ROOT
    so = os.path.join(ROOT, "oracle", "_ref", "libcfun_boxes_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.cfun_ref_nms.restype = ctypes.c_int
    g = load_golden("nms")

    def c_nms(b, s, thr, mx):
        b = np.ascontiguousarray(b, dtype=np.float32); s = np.ascontiguousarray(s, dtype=np.float32)
        keep = np.zeros(max(mx, 1), dtype=np.int32)
        n = lib.cfun_ref_nms(b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), b.shape[0],
                             ctypes.c_float(thr), mx, keep.ctypes.data_as(ctypes.c_void_p))
        return keep[:n]
    for case in ("rand_t7_m50", "rand_t3_all", "nested_degenerate", "integer_boxes_t3"):
        assert np.array_equal(c_nms(g[case + "/boxes"], g[case + "/scores"], float(g[case + "/thr"]), int(g[case + "/max"])), g[case + "/keep"])
    rng = np.random.default_rng(5)
    c = rng.uniform(0, 256, size=(4000, 3)); s = rng.uniform(16, 128, size=(4000, 3))
    b = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.float32)
    sc = rng.uniform(0, 1, size=4000).astype(np.float32)
    assert np.array_equal(c_nms(b, sc, 0.7, 500), O.non_max_suppression(b, sc, 0.7, 500))


def test_inference_pre_post_processing_matches_reference_golden():
    """mold_inputs ('self' resize + normalise) and unmold_detections (box rescale, trilinear unmold_mask, argmax) of the
    oracle against MaskRCNN.detect() of the unmodified reference (oracle/gen_golden_inference.py, BASELINE config 1)."""
    import scipy.ndimage as ndi
    g = load_golden("inference64")
    vol = g["vol"]
    z = ndi.zoom(vol.astype(np.float64), [64 / 80, 64 / 72, 64 / 48], order=1, mode="grid-constant", cval=0, grid_mode=True)
    assert np.array_equal(O.zoom_linear(vol, (64, 64, 64)), z)            # the restatement IS scipy's arithmetic, bit for bit
    molded, window = O.mold_inputs(vol[..., None], 64, 64)
    assert rel_err(molded, g["molded"]) < 1e-6 and tuple(window) == tuple(g["window"])
    boxes, scores, full = O.unmold_detections(g["detections"], g["mask0"], (1, 48, 80, 72), g["window"])
    assert np.array_equal(boxes, g["rois"]) and np.array_equal(scores, g["scores"])
    assert np.array_equal(full.astype(np.uint8), g["full_mask"])
