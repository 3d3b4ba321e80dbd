"""CFUN model (3-D Faster-R-CNN + U-Net mask head) on the cfun_b200 sm_100a CUDA ops.

Drop-in surface of reference model.py: the same module / function names, positional signatures, return structures and
state_dict keys (FPN:124, proposal_layer:199, RoI_Align:265, pyramid_roi_align:292, bbox_overlaps:377,
detection_target_layer:414, refine_detections:584, detection_layer:679, RPN:700, Classifier:750, Mask:787, the six
losses :808-981, compute_losses:984, build_rpn_targets:1090, MaskRCNN:1245, compose/parse_image_meta:1871/1891,
mold_image:1902).  What changed is where the work runs: every op on the hot path is a hand-written CUDA kernel behind
the C ABI (include/cfun_b200.h); box sorting / NMS / target assembly stay on the device instead of round-tripping
through numpy; masks targets are produced as class-index volumes instead of float64 one-hot stacks (see
detection_target_layer).  CUDA only -- there is no CPU path.
"""
import math
import os
import re
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import utils
from . import backbone
from . import mask_branch
from .layers import Conv3d, FrozenBatchNorm3d, Slot


def log(text, array=None):
    if array is not None:
        text = text.ljust(25)
        text += ("shape: {:20}  min: {:10.5f}  max: {:10.5f}".format(str(array.shape), array.min() if array.size else "",
                                                                     array.max() if array.size else ""))
    print(text)


def _f32(vals, device):
    return torch.tensor([float(v) for v in vals], dtype=torch.float32, device=device)


def compute_backbone_shapes(config, image_shape):
    """[N, (depth, height, width)] per pyramid level; image_shape is (H, W, D, C) (reference model.py:91-101)."""
    H, W, D = image_shape[:3]
    return np.array([[int(math.ceil(D / s)), int(math.ceil(H / s)), int(math.ceil(W / s))] for s in config.BACKBONE_STRIDES])


# ---------------------------------------------------------------------------------------------------------
#  FPN
# ---------------------------------------------------------------------------------------------------------
class FPN(nn.Module):
    def __init__(self, C1, C2, C3, out_channels, config):
        super().__init__()
        self.out_channels = out_channels
        self.C1, self.C2, self.C3 = C1, C2, C3
        self.P3_conv1 = Conv3d(config.BACKBONE_CHANNELS[1] * 4, out_channels, kernel_size=1, stride=1)
        self.P3_conv2 = Conv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.P2_conv1 = Conv3d(config.BACKBONE_CHANNELS[0] * 4, out_channels, kernel_size=1, stride=1)
        self.P2_conv2 = Conv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        c2 = self.C2(self.C1(x))
        c3 = self.C3(c2)
        p3 = self.P3_conv1(c3)
        p2 = self.P2_conv1(c2) + ops.upsample2x(p3)
        return [self.P2_conv2(p2), self.P3_conv2(p3)]


# ---------------------------------------------------------------------------------------------------------
#  Proposal layer
# ---------------------------------------------------------------------------------------------------------
def apply_box_deltas(boxes, deltas):
    """[N,6] boxes (z1,y1,x1,z2,y2,x2) refined by [N,6] deltas (reference model.py:155-182), on device."""
    n = boxes.shape[0]
    big = 3.0e38
    out, _ = ops.decode_clip(boxes.detach().float(), deltas.detach().float(), None, None, n, (1,) * 6,
                             (-big, -big, -big, big, big, big))
    return out


def clip_boxes(boxes, window):
    lo = _f32([window[0], window[1], window[2]] * 2, boxes.device)
    hi = _f32([window[3], window[4], window[5]] * 2, boxes.device)
    return torch.max(torch.min(boxes, hi), lo)


def proposal_layer(inputs, proposal_count, nms_threshold, anchors, config=None):
    """RPN outputs -> normalised proposals [1, n, 6] (reference model.py:199-258), entirely on device:
    bitonic sort of the fg scores, fused decode+clip of the top PRE_NMS_LIMIT anchors, bit-mask NMS, gather+normalise.
    One host sync (the kept count sizes the output, as the reference's dynamic shape does)."""
    probs = inputs[0].squeeze(0)
    deltas = inputs[1].squeeze(0)
    A = anchors.shape[0]
    scores = probs[:, 1].detach().contiguous()
    order = ops.sort_desc(scores)
    k = min(config.PRE_NMS_LIMIT, A)
    height, width, depth = [int(v) for v in config.IMAGE_SHAPE[:3]]
    boxes, _ = ops.decode_clip(anchors, deltas, None, order, k, config.RPN_BBOX_STD_DEV,
                               (0, 0, 0, depth, height, width))
    keep, count = ops.nms3d(boxes, nms_threshold, proposal_count)
    n = int(count.item())
    rois = ops.gather_boxes(boxes, keep, count, max(n, 1), (depth, height, width, depth, height, width))[:n]
    return rois.unsqueeze(0)


def proposal_layer_device(inputs, proposal_count, nms_threshold, anchors, config):
    """proposal_layer up to (not including) its read-back of the kept count: -> (boxes [k,6] in pixels and descending
    score order, keep int32 [proposal_count], count int32 [1]), all on the device.  The training step feeds these to
    detection_targets_begin, which normalises the kept boxes exactly as proposal_layer does."""
    probs = inputs[0].squeeze(0)
    deltas = inputs[1].squeeze(0)
    A = anchors.shape[0]
    scores = probs[:, 1].detach().contiguous()
    order = ops.sort_desc(scores)
    k = min(config.PRE_NMS_LIMIT, A)
    height, width, depth = [int(v) for v in config.IMAGE_SHAPE[:3]]
    boxes, _ = ops.decode_clip(anchors, deltas, None, order, k, config.RPN_BBOX_STD_DEV,
                               (0, 0, 0, depth, height, width))
    keep, count = ops.nms3d(boxes, nms_threshold, proposal_count)
    return boxes, keep, count


# ---------------------------------------------------------------------------------------------------------
#  RoI crop-resize ("RoIAlign")
# ---------------------------------------------------------------------------------------------------------
def RoI_Align(feature_map, pool_size, boxes, out_ncdhw=False):
    """feature_map [C,D,H,W]; boxes [n,6] normalised -> [n,C,*pool_size] (reference model.py:265-289)."""
    fm = feature_map.unsqueeze(0)
    return ops.roi_crop_resize(fm, None, boxes, None, pool_size, out_ncdhw)


def log2(x):
    return torch.log(x) / math.log(2.0)


def pyramid_roi_align(inputs, pool_size, test_flag=False, out_ncdhw=False):
    """inputs = [boxes, P2, P3]; one kernel launch covers every box of both levels, output rows stay in box order
    (the reference gathers per level and sorts back, model.py:334-368).  Gradients flow to the feature maps only."""
    boxes = inputs[0]
    if boxes.dim() == 3:
        boxes = boxes.squeeze(0)
    maps = [m if m.dim() == 5 else m.unsqueeze(0) for m in inputs[1:]]
    boxes = boxes.detach()
    level = ops.roi_level(boxes)
    return ops.roi_crop_resize(maps[0], maps[1], boxes, level, pool_size, out_ncdhw)


# ---------------------------------------------------------------------------------------------------------
#  Detection targets
# ---------------------------------------------------------------------------------------------------------
def bbox_overlaps(boxes1, boxes2):
    return ops.bbox_overlaps3d(boxes1, boxes2)


def _label_volume(gt_masks):
    """Accepts the reference's one-hot stack [C,D,H,W] (float) or a class-id volume [D,H,W] (int) -> int32 [D,H,W]."""
    if gt_masks.dim() == 3:
        return gt_masks.to(torch.int32).contiguous()
    return torch.argmax(gt_masks, dim=0).to(torch.int32).contiguous()


def detection_target_layer(proposals, gt_class_ids, gt_boxes, gt_masks, config):
    """Sub-samples proposals and builds class / box / mask targets (reference model.py:414-563).

    Same RNG contract as the reference: the two sub-sampling permutations are torch.randperm draws on the host
    generator, in the same order.  Returns (positive_rois, rois, class_ids, deltas, masks).  `masks` is an int64
    class-index volume [P, *MASK_SHAPE] (argmax of the reference's float64 one-hot [P,8,*MASK_SHAPE]); set
    config.DENSE_MASK_TARGETS = True to get the one-hot float64 stack itself.  The losses accept both."""
    proposals = proposals.squeeze(0) if proposals.dim() == 3 else proposals
    gt_class_ids = gt_class_ids.squeeze(0) if gt_class_ids.dim() == 2 else gt_class_ids
    gt_boxes = gt_boxes.squeeze(0) if gt_boxes.dim() == 3 else gt_boxes
    if gt_masks.dim() in (5, 4) and gt_masks.shape[0] == 1:
        gt_masks = gt_masks.squeeze(0)
    dev = proposals.device
    empty = torch.zeros((0, 6), device=dev)
    if proposals.shape[0] == 0:
        return empty, empty, torch.zeros(0, dtype=torch.long, device=dev), empty, torch.zeros(0, device=dev)

    overlaps = bbox_overlaps(proposals, gt_boxes)
    roi_iou_max = overlaps.max(dim=1)[0]
    thr = config.DETECTION_TARGET_IOU_THRESHOLD
    pos_all = torch.nonzero(roi_iou_max >= thr)[:, 0]
    neg_all = torch.nonzero(roi_iou_max < thr)[:, 0]

    positive_count = 0
    if pos_all.numel() > 0:
        want = config.TRAIN_ROIS_PER_IMAGE * config.ROI_POSITIVE_RATIO
        want = int(round(want)) if getattr(config, "ROI_COUNT_ROUND", False) else int(want)      # LiTS_2017/model.py:448 rounds
        perm = torch.randperm(pos_all.numel())[:want].to(dev)
        positive_indices = pos_all[perm]
        positive_count = positive_indices.numel()
        positive_rois = proposals[positive_indices]
        assign = overlaps[positive_indices].max(dim=1)[1]
        roi_gt_boxes = gt_boxes[assign]
        roi_gt_class_ids = gt_class_ids[assign]
        deltas = ops.box_refinement(positive_rois, roi_gt_boxes, config.BBOX_STD_DEV)
        label = _label_volume(gt_masks)
        dense = bool(getattr(config, "DENSE_MASK_TARGETS", False))
        onehot, index = ops.mask_target_crop(label, positive_rois, getattr(config, "NUM_CLASSES", 8) if gt_masks.dim() == 3 else gt_masks.shape[0],
                                             config.MASK_SHAPE, onehot=dense, index=not dense)
        masks = onehot if dense else index

    negative_count = 0
    if neg_all.numel() > 0 and positive_count > 0:
        want = (1.0 / config.ROI_POSITIVE_RATIO) * positive_count - positive_count
        want = int(round(want)) if getattr(config, "ROI_COUNT_ROUND", False) else int(want)
        perm = torch.randperm(neg_all.numel())[:want].to(dev)
        negative_indices = neg_all[perm]
        negative_count = negative_indices.numel()
        negative_rois = proposals[negative_indices]

    if positive_count > 0 and negative_count > 0:
        rois = torch.cat((positive_rois, negative_rois), dim=0)
        class_ids = torch.cat([roi_gt_class_ids.long(), torch.zeros(negative_count, dtype=torch.long, device=dev)], dim=0)
        deltas = torch.cat([deltas, torch.zeros((negative_count, 6), device=dev)], dim=0)
        return positive_rois, rois, class_ids, deltas, masks
    if positive_count > 0:
        return positive_rois, positive_rois, roi_gt_class_ids.long(), deltas, masks
    return empty, empty, torch.zeros(0, dtype=torch.long, device=dev), empty, torch.zeros(0, device=dev)


def detection_targets_begin(boxes, keep, count, gt_boxes, config, extra_counts=None):
    """First half of detection_target_layer for the training step, with no host round trip: one kernel normalises the
    kept proposals, takes max / argmax IoU against the ground-truth boxes and compacts the positive / negative index
    lists (ops.roi_candidates); the three counts (+ extra_counts, any int32 device vector the caller wants read back at
    the same time) start their copy to pinned host memory.  Returns the state detection_targets_finish consumes."""
    height, width, depth = [int(v) for v in config.IMAGE_SHAPE[:3]]
    gt_boxes = gt_boxes.squeeze(0) if gt_boxes.dim() == 3 else gt_boxes
    rois, assign, pos_list, neg_list, counts = ops.roi_candidates(boxes, keep, count, (depth, height, width, depth, height, width),
                                                                  gt_boxes, config.DETECTION_TARGET_IOU_THRESHOLD)
    if extra_counts is not None:
        counts = torch.cat([counts, extra_counts.to(torch.int32)])
    host = torch.empty(counts.shape[0], dtype=torch.int32, pin_memory=True)
    host.copy_(counts, non_blocking=True)
    ready = torch.cuda.Event()
    ready.record()
    return dict(rois=rois, assign=assign, pos_list=pos_list, neg_list=neg_list, host=host, ready=ready, gt_boxes=gt_boxes)


def detection_targets_finish(state, gt_class_ids, gt_masks, config):
    """Second half: the step's ONE wait on the device (the counts), the two torch.randperm draws on the host generator in the
    reference's order (model.py:441-444, 523-526), then one kernel for the sampled RoIs, class ids and box deltas
    (ops.roi_targets) and one for the mask targets.  Same return values as detection_target_layer (the reference-shaped
    function above, which the parity tests compare this against); also returns the extra counts read back."""
    state["ready"].synchronize()
    host = state["host"].tolist()
    n, n_pos, n_neg = host[:3]
    extra = host[3:]
    rois_all = state["rois"]
    dev = rois_all.device
    gt_class_ids = gt_class_ids.squeeze(0) if gt_class_ids.dim() == 2 else gt_class_ids
    if gt_masks.dim() in (5, 4) and gt_masks.shape[0] == 1:
        gt_masks = gt_masks.squeeze(0)
    empty = torch.zeros((0, 6), device=dev)
    none = (empty, empty, torch.zeros(0, dtype=torch.long, device=dev), empty, torch.zeros(0, device=dev))
    if n == 0 or n_pos == 0:
        return none, extra
    rnd = bool(getattr(config, "ROI_COUNT_ROUND", False))
    want = config.TRAIN_ROIS_PER_IMAGE * config.ROI_POSITIVE_RATIO
    want = int(round(want)) if rnd else int(want)
    perm_p = torch.randperm(n_pos)[:want]
    P = int(perm_p.numel())
    if P == 0:
        return none, extra
    perm_n = None
    if n_neg > 0:
        want = (1.0 / config.ROI_POSITIVE_RATIO) * P - P
        want = int(round(want)) if rnd else int(want)
        perm_n = torch.randperm(n_neg)[:want]
    Rn = int(perm_n.numel()) if perm_n is not None else 0
    R = P + Rn
    perm_host = torch.empty(R, dtype=torch.int64, pin_memory=True)
    perm_host[:P] = perm_p
    if Rn:
        perm_host[P:] = perm_n
    perm = perm_host.to(dev, non_blocking=True)
    rois, class_ids, deltas = ops.roi_targets(rois_all, state["assign"], state["pos_list"], state["neg_list"], perm, P, R,
                                              state["gt_boxes"], gt_class_ids, config.BBOX_STD_DEV)
    positive_rois = rois[:P]
    label = _label_volume(gt_masks)
    dense = bool(getattr(config, "DENSE_MASK_TARGETS", False))
    onehot, index = ops.mask_target_crop(label, positive_rois, getattr(config, "NUM_CLASSES", 8) if gt_masks.dim() == 3 else gt_masks.shape[0],
                                         config.MASK_SHAPE, onehot=dense, index=not dense)
    masks = onehot if dense else index
    return (positive_rois, rois, class_ids, deltas, masks), extra


# ---------------------------------------------------------------------------------------------------------
#  Detection layer (inference)
# ---------------------------------------------------------------------------------------------------------
def clip_to_window(window, boxes):
    return clip_boxes(boxes, window)


def refine_detections(rois, probs, deltas, window, config):
    """[N,6] rois + class probs + class deltas -> [n,(z1,y1,x1,z2,y2,x2,class_id,score)] (reference model.py:584-676),
    including its use of RPN_BBOX_STD_DEV for the head deltas (:610)."""
    dev = rois.device
    class_ids = probs.argmax(dim=1)
    idx = torch.arange(class_ids.shape[0], device=dev)
    class_scores = probs[idx, class_ids]
    deltas_specific = deltas[idx, class_ids]
    refined = apply_box_deltas(rois, deltas_specific * _f32(config.RPN_BBOX_STD_DEV, dev).view(1, 6))
    height, width, depth = [int(v) for v in config.IMAGE_SHAPE[:3]]
    refined = refined * _f32([depth, height, width, depth, height, width], dev)
    refined = torch.round(clip_to_window([float(w) for w in window], refined))
    keep_bool = class_ids > 0
    if config.DETECTION_MIN_CONFIDENCE:
        keep_bool = keep_bool & (class_scores >= config.DETECTION_MIN_CONFIDENCE)
    keep = torch.nonzero(keep_bool)[:, 0]
    if keep.numel() == 0:
        # the reference dies here with UnboundLocalError (model.py:641-662); report it as what it is
        raise RuntimeError("refine_detections: no RoI passes class>0 and score>=DETECTION_MIN_CONFIDENCE")
    pre_cls, pre_scores, pre_rois = class_ids[keep], class_scores[keep], refined[keep]
    nms_keep = []
    for cid in torch.unique(pre_cls):
        ixs = torch.nonzero(pre_cls == cid)[:, 0]
        order = ops.sort_desc(pre_scores[ixs]).long()
        kk, cnt = ops.nms3d(pre_rois[ixs][order], config.DETECTION_NMS_THRESHOLD, config.DETECTION_MAX_INSTANCES)
        ck = kk[:int(cnt.item())].long()
        nms_keep.append(keep[ixs[order[ck]]])
    nms_keep = torch.unique(torch.cat(nms_keep))
    mask = torch.zeros(class_ids.shape[0], dtype=torch.bool, device=dev)
    mask[nms_keep] = True
    keep = keep[mask[keep]]
    roi_count = min(config.DETECTION_MAX_INSTANCES, keep.numel())
    top = ops.sort_desc(class_scores[keep])[:roi_count].long()
    keep = keep[top]
    return torch.cat((refined[keep], class_ids[keep].unsqueeze(1).float(), class_scores[keep].unsqueeze(1)), dim=1)


def detection_layer(config, rois, mrcnn_class, mrcnn_bbox, image_meta):
    rois = rois.squeeze(0)
    _, _, window, _ = parse_image_meta(image_meta)
    return refine_detections(rois, mrcnn_class, mrcnn_bbox, window[0], config)


# ---------------------------------------------------------------------------------------------------------
#  RPN and heads
# ---------------------------------------------------------------------------------------------------------
class RPN(nn.Module):
    """reference model.py:700-743: 3x3x3 shared conv + ReLU (fused epilogue), two 1x1x1 heads, softmax over (bg, fg).
    The NDHWC conv output *is* the reference's permute(0,2,3,4,1).contiguous().view(B,-1,K) layout, so the reshape is free."""

    def __init__(self, anchors_per_location, anchor_stride, channel, conv_channel):
        super().__init__()
        self.conv_shared = Conv3d(channel, conv_channel, kernel_size=3, stride=anchor_stride, padding=1)
        self.relu = Slot("ReLU (conv epilogue)")
        self.conv_class = Conv3d(conv_channel, 2 * anchors_per_location, kernel_size=1, stride=1)
        self.softmax = nn.Softmax(dim=2)
        self.conv_bbox = Conv3d(conv_channel, 6 * anchors_per_location, kernel_size=1, stride=1)

    def forward(self, x):
        x = self.conv_shared(x, relu=True)
        B = x.shape[0]
        rpn_class_logits = self.conv_class(x).permute(0, 2, 3, 4, 1).reshape(B, -1, 2)
        rpn_probs = self.softmax(rpn_class_logits)
        rpn_bbox = self.conv_bbox(x).permute(0, 2, 3, 4, 1).reshape(B, -1, 6)
        return [rpn_class_logits, rpn_probs, rpn_bbox]


class Classifier(nn.Module):
    """reference model.py:750-784.  conv1 has kernel == pool size: a 221184 -> fc_size product, weight-bandwidth bound."""

    def __init__(self, channel, pool_size, image_shape, num_classes, fc_size, test_flag=False):
        super().__init__()
        self.pool_size = pool_size
        self.image_shape = image_shape
        self.fc_size = fc_size
        self.test_flag = test_flag
        self.conv1 = Conv3d(channel, fc_size, kernel_size=tuple(pool_size), stride=1)
        self.bn1 = FrozenBatchNorm3d(fc_size, eps=0.001, momentum=0.01)
        self.conv2 = Conv3d(fc_size, fc_size, kernel_size=1, stride=1)
        self.bn2 = FrozenBatchNorm3d(fc_size, eps=0.001, momentum=0.01)
        self.relu = Slot("ReLU (fused into the BN pass)")
        self.linear_class = nn.Linear(fc_size, num_classes)
        self.softmax = nn.Softmax(dim=1)
        self.linear_bbox = nn.Linear(fc_size, num_classes * 6)

    def forward(self, x, rois):
        x = pyramid_roi_align([rois] + x, self.pool_size, self.test_flag, out_ncdhw=True)
        x = ops.fc_conv(x, self.conv1.weight, self.conv1.bias)
        x = self.bn1(x, relu=True)
        x = self.bn2(self.conv2(x), relu=True)
        x = x.reshape(-1, self.fc_size)
        mrcnn_class_logits = self.linear_class(x)
        mrcnn_probs = self.softmax(mrcnn_class_logits)
        mrcnn_bbox = self.linear_bbox(x)
        return [mrcnn_class_logits, mrcnn_probs, mrcnn_bbox.view(mrcnn_bbox.shape[0], -1, 6)]


class Mask(nn.Module):
    """reference model.py:787-801: crops of the *input volume* (C=1) -> Modified3DUNet -> softmax over classes."""

    def __init__(self, channel, pool_size, num_classes, conv_channel, stage, test_flag=False):
        super().__init__()
        self.pool_size = pool_size
        self.test_flag = test_flag
        self.modified_u_net = mask_branch.Modified3DUNet(channel, num_classes, stage, conv_channel)
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x, rois):
        x = pyramid_roi_align([rois] + x, self.pool_size, self.test_flag)
        x = self.modified_u_net(x)
        return x, self.softmax(x)


# ---------------------------------------------------------------------------------------------------------
#  Losses
# ---------------------------------------------------------------------------------------------------------
def compute_rpn_class_loss(rpn_match, rpn_class_logits):
    rpn_match = rpn_match.squeeze(2)
    anchor_class = (rpn_match == 1).long()
    ind = torch.nonzero(rpn_match != 0)
    return F.cross_entropy(rpn_class_logits[ind[:, 0], ind[:, 1], :], anchor_class[ind[:, 0], ind[:, 1]])


def compute_rpn_bbox_loss(target_bbox, rpn_match, rpn_bbox):
    rpn_match = rpn_match.squeeze(2)
    ind = torch.nonzero(rpn_match == 1)
    rpn_bbox = rpn_bbox[ind[:, 0], ind[:, 1]]
    return F.smooth_l1_loss(rpn_bbox, target_bbox[0, :rpn_bbox.shape[0], :])


def _nonzero_static(mask, size):
    """ascending indices of the True entries of a 1-D mask, padded with -1 to `size`, without a host round trip"""
    try:
        return torch.nonzero_static(mask, size=size, fill_value=-1)[:, 0]
    except (AttributeError, NotImplementedError, RuntimeError):
        n = mask.shape[0]
        pos = torch.cumsum(mask, 0) - 1
        out = torch.full((size + 1,), -1, dtype=torch.long, device=mask.device)
        slot = torch.where(mask, pos.clamp(max=size), torch.full_like(pos, size))
        out.scatter_(0, slot, torch.arange(n, device=mask.device))
        out[size] = -1
        return out[:size]


def compute_rpn_losses_static(rpn_match, target_bbox, rpn_class_logits, rpn_pred_bbox, max_anchors):
    """compute_rpn_class_loss + compute_rpn_bbox_loss (reference model.py:836-873) for batch 1 without their torch.nonzero
    read-backs: the <= max_anchors non-neutral anchors (RPN_TRAIN_ANCHORS_PER_IMAGE, the cap build_rpn_targets enforces)
    are gathered into fixed-size index lists, padded rows carry weight 0.  Also returns the int32 device vector
    [non-neutral, positive] so that the caller can verify the cap when it next reads from the device."""
    m = rpn_match.reshape(-1)
    valid, pos = m != 0, m == 1
    counts = torch.stack([valid.sum(), pos.sum()]).to(torch.int32)
    ind = _nonzero_static(valid, max_anchors)
    w = (ind >= 0).to(rpn_class_logits.dtype)
    idx = ind.clamp(min=0)
    ce = F.cross_entropy(rpn_class_logits[0].index_select(0, idx), (m.index_select(0, idx) == 1).long(), reduction='none')
    class_loss = (ce * w).sum() / w.sum()
    T = min(max_anchors, target_bbox.shape[1])
    pind = _nonzero_static(pos, T)
    pw = (pind >= 0).to(rpn_pred_bbox.dtype)
    pred = rpn_pred_bbox[0].index_select(0, pind.clamp(min=0))
    l1 = F.smooth_l1_loss(pred, target_bbox[0, :T, :], reduction='none') * pw[:, None]
    bbox_loss = l1.sum() / (pw.sum() * pred.shape[1])
    return class_loss, bbox_loss, counts


def _zero_loss(like):
    return torch.zeros(1, device=like.device)


def compute_mrcnn_class_loss(target_class_ids, pred_class_logits):
    if target_class_ids.shape[0] == 0:
        return _zero_loss(target_class_ids)
    return F.cross_entropy(pred_class_logits, target_class_ids.long())


def compute_mrcnn_bbox_loss(target_bbox, target_class_ids, pred_bbox):
    if target_class_ids.shape[0] == 0:
        return _zero_loss(target_class_ids)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    cls = target_class_ids[pos].long()
    return F.smooth_l1_loss(pred_bbox[pos, cls, :], target_bbox[pos, :])


def _mask_index(target_masks, rows):
    """class-index volume [P,d,h,w] from either target representation"""
    if target_masks.dim() == 5:
        return torch.argmax(target_masks[rows].long(), dim=1)
    return target_masks[rows]


def compute_mrcnn_mask_loss(target_masks, target_class_ids, pred_masks, class_weight=None):
    """CrossEntropy between the mask logits [P,ncls,d,h,w] and the argmax of the one-hot target (reference :909-935)."""
    if target_class_ids.shape[0] == 0:
        return _zero_loss(target_class_ids)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    y_true = _mask_index(target_masks, pos)
    return ops.mask_cross_entropy(pred_masks[pos], y_true, class_weight)


def compute_mrcnn_mask_edge_loss(target_masks, target_class_ids, pred_masks, mode="magnitude"):
    """3-D Sobel edge-agreement loss (reference :938-981; mode "raw" = LiTS_2017/model.py:943-981) as one fused stencil
    kernel (forward) + two (backward)."""
    if target_class_ids.shape[0] == 0:
        return _zero_loss(target_class_ids)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    P = pos.shape[0]
    tgt = _mask_index(target_masks, torch.arange(P, device=pos.device))     # reference takes target_masks[:P]
    return ops.sobel_edge_loss(pred_masks[pos], tgt, ops.SOBEL_RAW if mode == "raw" else ops.SOBEL_MAGNITUDE)


def compute_losses(rpn_match, rpn_bbox, rpn_class_logits, rpn_pred_bbox, target_class_ids, mrcnn_class_logits,
                   target_deltas, mrcnn_bbox, target_mask, mrcnn_mask, mrcnn_mask_logits, stage, config=None):
    if config is not None and getattr(config, "STAGED_LOSSES", False):
        # LiTS_2017/model.py:985-1001: detector losses in 'beginning', mask losses (weighted CE + raw-Sobel edge) afterwards
        zero = _zero_loss(rpn_class_logits)
        binary_ids = (target_class_ids > 0).long()
        if stage == 'beginning':
            return [compute_rpn_class_loss(rpn_match, rpn_class_logits), compute_rpn_bbox_loss(rpn_bbox, rpn_match, rpn_pred_bbox),
                    compute_mrcnn_class_loss(binary_ids, mrcnn_class_logits),
                    compute_mrcnn_bbox_loss(target_deltas, binary_ids, mrcnn_bbox), zero, zero]
        w = getattr(config, "MASK_CLASS_WEIGHT", None)
        cw = None if w is None else torch.tensor(w, dtype=torch.float32, device=mrcnn_mask_logits.device)
        return [zero, zero, zero, zero, compute_mrcnn_mask_loss(target_mask, target_class_ids, mrcnn_mask_logits, cw),
                compute_mrcnn_mask_edge_loss(target_mask, target_class_ids, mrcnn_mask, config.EDGE_LOSS_MODE)]
    rpn_class_loss = compute_rpn_class_loss(rpn_match, rpn_class_logits)
    rpn_bbox_loss = compute_rpn_bbox_loss(rpn_bbox, rpn_match, rpn_pred_bbox)
    binary_ids = (target_class_ids > 0).long()
    mrcnn_class_loss = compute_mrcnn_class_loss(binary_ids, mrcnn_class_logits)
    mrcnn_bbox_loss = compute_mrcnn_bbox_loss(target_deltas, binary_ids, mrcnn_bbox)
    mrcnn_mask_loss = compute_mrcnn_mask_loss(target_mask, target_class_ids, mrcnn_mask_logits)
    if stage == 'finetune':
        mrcnn_mask_edge_loss = compute_mrcnn_mask_edge_loss(target_mask, target_class_ids, mrcnn_mask)
    else:
        mrcnn_mask_edge_loss = _zero_loss(rpn_class_logits)
    return [rpn_class_loss, rpn_bbox_loss, mrcnn_class_loss, mrcnn_bbox_loss, mrcnn_mask_loss, mrcnn_mask_edge_loss]


# ---------------------------------------------------------------------------------------------------------
#  Host-side data preparation (off the hot path; kept so train_model / detect can run)
# ---------------------------------------------------------------------------------------------------------
from .workload import build_rpn_targets      # host-side numpy restatement (reference model.py:1090-1181), shared with bench.py


def mold_image(images):
    """(x - mean) / std (reference model.py:1902-1904); the device kernel for int16 volumes is ops.mold_volume_i16."""
    return (images - images.mean()) / images.std()


def compose_image_meta(image_id, image_shape, window, active_class_ids):
    return np.array([image_id] + list(image_shape) + list(window) + list(active_class_ids))


def parse_image_meta(meta):
    return meta[:, 0], meta[:, 1:5], meta[:, 5:11], meta[:, 11:]


def load_image_gt(image, mask, angle, dataset, config, anchors):
    """Ground truth for one volume (reference model.py:1007-1087).  image [H,W,D,1], mask [H,W,D] class ids.
    The in-plane rotation augmentation uses scipy.ndimage (nearest neighbour) in place of imgaug."""
    if angle:
        import scipy.ndimage as ndi
        image = ndi.rotate(image, angle, axes=(0, 1), reshape=False, order=0, mode="constant", cval=0)
        mask = ndi.rotate(mask.astype(np.uint8), angle, axes=(0, 1), reshape=False, order=0, mode="constant", cval=0)
    mask = mask.astype(np.int32)
    image = image.transpose((3, 2, 0, 1))
    mask = mask.transpose((2, 0, 1))
    bbox = utils.extract_bboxes(np.expand_dims(mask, -1)).astype(np.float64)
    z1, y1, x1, z2, y2, x2 = bbox[0]
    d, h, w = z2 - z1, y2 - y1, x2 - x1
    lo = np.floor(np.maximum(0, [z1 - d * 0.05, y1 - h * 0.05, x1 - w * 0.05]))
    hi = np.ceil(np.minimum(mask.shape, [z2 + d * 0.05, y2 + h * 0.05, x2 + w * 0.05]))
    bbox = np.tile(np.concatenate([lo, hi]).astype(np.int32)[None], (config.NUM_CLASSES - 1, 1))
    masks, class_ids = dataset.process_mask(mask)
    rpn_match, rpn_bbox = build_rpn_targets(anchors, np.array([bbox[0]]), config)
    return mold_image(image.astype(np.float32)), rpn_match[:, np.newaxis], rpn_bbox, class_ids, bbox, masks


class Dataset(torch.utils.data.Dataset):
    """reference model.py:1184-1238"""

    def __init__(self, dataset, config):
        self.image_ids = np.copy(dataset.image_ids)
        self.dataset = dataset
        self.config = config

    def __getitem__(self, image_index):
        image_id = self.image_ids[image_index]
        image = self.dataset.load_image(image_id)
        mask = self.dataset.load_mask(image_id)
        c = self.config
        image, window, scale, padding, crop = utils.resize_image(image, min_dim=c.IMAGE_MIN_DIM, max_dim=c.IMAGE_MAX_DIM,
                                                                 min_scale=c.IMAGE_MIN_SCALE, mode=c.IMAGE_RESIZE_MODE)
        mask = utils.resize_mask(mask, scale, padding, max_dim=c.IMAGE_MAX_DIM, min_dim=c.IMAGE_MIN_DIM, crop=crop,
                                 mode=c.IMAGE_RESIZE_MODE)
        active = np.zeros([self.dataset.num_classes], dtype=np.int32)
        active[self.dataset.source_class_ids[self.dataset.image_info[image_id]["source"]]] = 1
        return image, compose_image_meta(image_id, image.shape, window, active), mask

    def __len__(self):
        return self.image_ids.shape[0]


# ---------------------------------------------------------------------------------------------------------
#  Heads + head losses as one static-shape callable (CUDA-graph capturable: no host syncs, no dynamic shapes)
# ---------------------------------------------------------------------------------------------------------
class PendingLosses:
    """The 7 loss scalars of one step on their way to pinned host memory (train_step_from_host(wait=False))."""

    def __init__(self, dev_tensor):
        self.host = torch.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True)
        self.host.copy_(dev_tensor, non_blocking=True)
        self.ready = torch.cuda.Event()
        self.ready.record()

    def result(self):
        self.ready.synchronize()
        return self.host


class HeadsTail(nn.Module):
    """classifier head + mask head + their four losses for a fixed (P positives, R RoIs) split, positives first
    (the layout detection_target_layer produces).  Numerically the same ops as Classifier / Mask / compute_*_loss; the
    only difference is that the positive rows are addressed as [:P] instead of through torch.nonzero, which lets
    torch.cuda.make_graphed_callables capture forward and backward (about 1500 of the step's kernel launches)."""

    def __init__(self, classifier, mask, stage, config=None):
        super().__init__()
        self.classifier = classifier
        self.mask = mask
        self.stage = stage
        self.staged = bool(getattr(config, "STAGED_LOSSES", False))
        self.edge_mode = ops.SOBEL_RAW if getattr(config, "EDGE_LOSS_MODE", "magnitude") == "raw" else ops.SOBEL_MAGNITUDE
        w = getattr(config, "MASK_CLASS_WEIGHT", None)
        self.class_weight = None if w is None else [float(v) for v in w]
        self.class_weight_t = None          # device tensor, created eagerly (never inside a CUDA-graph capture): ensure_device()

    def ensure_device(self, dev):
        if self.class_weight is not None and (self.class_weight_t is None or self.class_weight_t.device != dev):
            self.class_weight_t = torch.tensor(self.class_weight, dtype=torch.float32, device=dev)
        return self

    def forward(self, p2, p3, image, rois, p_rois, class_ids, deltas, mask_index, d0, d1, d2, d3, d4):
        P = p_rois.shape[0]
        unet = self.mask.modified_u_net
        saved = unet.injected_drop
        unet.injected_drop = [d0, d1, d2, d3, d4] if unet.training and unet.use_dropout else None
        zero = torch.zeros((), device=p2.device)
        run_cls = not (self.staged and self.stage != 'beginning')       # LiTS: the heads of the stage that is not trained are skipped
        run_mask = not (self.staged and self.stage == 'beginning')
        try:
            if run_cls:
                c_logits, _, c_bbox = self.classifier([p2, p3], rois)
            need_probs = self.stage == 'finetune' or (self.staged and self.stage != 'beginning')     # only the edge loss reads them
            if run_mask:
                if need_probs:
                    m_logits, m_probs = self.mask([image, image], p_rois)
                else:       # Mask.forward without its softmax (model.py:796-801): 2 passes over the 113 MB logits nobody reads
                    m_logits = unet(pyramid_roi_align([p_rois, image, image], self.mask.pool_size, self.mask.test_flag))
        finally:
            unet.injected_drop = saved
        l_cls = l_box = l_mask = l_edge = zero
        if run_cls:
            binary = (class_ids > 0).long()
            l_cls = F.cross_entropy(c_logits, binary)
            l_box = F.smooth_l1_loss(c_bbox[:P, 1, :], deltas[:P])
        if run_mask:
            l_mask = ops.mask_cross_entropy(m_logits, mask_index, self.class_weight_t)
            if self.stage == 'finetune' or (self.staged and self.stage != 'beginning'):
                l_edge = ops.sobel_edge_loss(m_probs, mask_index, self.edge_mode).reshape(())
        return torch.stack([l_cls, l_box, l_mask, l_edge])


# ---------------------------------------------------------------------------------------------------------
#  MaskRCNN
# ---------------------------------------------------------------------------------------------------------
LOSS_NAMES = ["rpn_class_loss", "rpn_bbox_loss", "mrcnn_class_loss", "mrcnn_bbox_loss", "mrcnn_mask_loss",
              "mrcnn_mask_edge_loss"]


class MaskRCNN(nn.Module):
    """Same constructor, attributes, state_dict (220 entries at HeartConfig widths) and methods as reference
    model.py:1245-1864."""

    def __init__(self, config, model_dir, test_flag=False):
        super().__init__()
        self.epoch = 0
        self.config = config
        self.model_dir = model_dir
        self.build(config=config, test_flag=test_flag)
        self.initialize_weights()
        self._graphed_tails = None
        self._roi_scale = None         # [d,h,w,d,h,w] on the device (proposal / ground-truth box normalisation)
        self._unit_drops = None        # all-ones Dropout3d masks for eval-mode / LiTS heads tails
        self._graph_ws_gen = 0
        self.graph_kernel_counts = {}
        self.graph_replays = {}

    def build(self, config, test_flag=False):
        if getattr(config, "TRAIN_BN", False):
            # the reference itself never trains BatchNorm on this path (model.py:1297-1304 freezes it, :1401-1406 forces eval);
            # FrozenBatchNorm3d has no statistics update and no scale/shift gradient, so refuse instead of silently ignoring
            raise NotImplementedError("config.TRAIN_BN = True is not supported: BatchNorm runs frozen (reference model.py:1297-1304)")
        h, w, d = config.IMAGE_SHAPE[:3]
        if h / 16 != int(h / 16) or w / 16 != int(w / 16) or d / 16 != int(d / 16):
            raise Exception("Image size must be dividable by 16. Use 256, 320, 512, ... etc.")
        factory = getattr(backbone, getattr(config, "BACKBONE", "P3D19"), backbone.P3D19)
        C1, C2, C3 = factory(config=config).stages()
        self.fpn = FPN(C1, C2, C3, out_channels=config.TOP_DOWN_PYRAMID_SIZE, config=config)
        anchors = utils.generate_pyramid_anchors(config.RPN_ANCHOR_SCALES, config.RPN_ANCHOR_RATIOS,
                                                 compute_backbone_shapes(config, config.IMAGE_SHAPE),
                                                 config.BACKBONE_STRIDES, config.RPN_ANCHOR_STRIDE)
        self.anchors = torch.from_numpy(anchors).float()
        self.rpn = RPN(len(config.RPN_ANCHOR_RATIOS), config.RPN_ANCHOR_STRIDE, config.TOP_DOWN_PYRAMID_SIZE,
                       config.RPN_CONV_CHANNELS)
        self.classifier = Classifier(config.TOP_DOWN_PYRAMID_SIZE, config.POOL_SIZE, config.IMAGE_SHAPE, 2,
                                     config.FPN_CLASSIFY_FC_LAYERS_SIZE, test_flag)
        self.mask = Mask(1, config.MASK_POOL_SIZE, config.NUM_CLASSES, config.UNET_MASK_BRANCH_CHANNEL, config.STAGE,
                         test_flag)
        # not registered as a sub-module (would duplicate state_dict keys): shares classifier / mask by reference
        self.mask.modified_u_net.use_dropout = bool(getattr(config, "UNET_DROPOUT", True))
        object.__setattr__(self, "_tail", HeadsTail(self.classifier, self.mask, config.STAGE, config))
        if getattr(config, "STAGED_LOSSES", False) and config.STAGE != 'beginning':
            for m in (self.fpn, self.rpn):          # LiTS_2017/model.py:1309-1311: the detector is frozen after 'beginning'
                for p in m.parameters():
                    p.requires_grad = False
        if not config.TRAIN_BN:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm3d):
                    for p in m.parameters():
                        p.requires_grad = False

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.anchors = fn(self.anchors)      # plain attribute in the reference (model.py:1276-1284): follow .cuda()
        return out

    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def set_trainable(self, layer_regex):
        for name, p in self.named_parameters():
            if not bool(re.fullmatch(layer_regex, name)):
                p.requires_grad = False

    def load_weights(self, file_path):
        if os.path.exists(file_path):
            self.load_state_dict(torch.load(file_path), strict=True)
            print("Weight file loading success!")
        else:
            print("Weight file not found ...")

    # -- forward ------------------------------------------------------------------------------------------
    def predict(self, inputs, mode):
        molded_images = inputs[0]
        image_metas = inputs[1]
        if not molded_images.is_cuda:
            raise RuntimeError("cfun_b200.MaskRCNN runs on CUDA (sm_100a) only; call .cuda() on the model and inputs")
        if mode == 'inference':
            self.eval()
        elif mode == 'training':
            self.train()     # BatchNorm stays frozen/eval regardless (FrozenBatchNorm3d)
        cfg = self.config

        p2_out, p3_out, rpn_class_logits, rpn_class, rpn_bbox, rpn_rois = self.rpn_proposals(molded_images, mode)
        mrcnn_classifier_feature_maps = [p2_out, p3_out]
        mrcnn_mask_feature_maps = [molded_images, molded_images]
        dev = molded_images.device
        h, w, d = cfg.IMAGE_SHAPE[:3]
        scale = _f32([d, h, w, d, h, w], dev)

        if mode == 'inference':
            mrcnn_class_logits, mrcnn_class, mrcnn_bbox = self.classifier(mrcnn_classifier_feature_maps, rpn_rois)
            detections = detection_layer(cfg, rpn_rois, mrcnn_class, mrcnn_bbox, image_metas)
            detection_boxes = (detections[:, :6] / scale).unsqueeze(0)
            _, mrcnn_mask = self.mask(mrcnn_mask_feature_maps, detection_boxes)
            return [detections.unsqueeze(0), mrcnn_mask.unsqueeze(0)]

        gt_class_ids, gt_boxes, gt_masks = inputs[2], inputs[3], inputs[4]
        gt_boxes = gt_boxes / scale
        p_rois, rois, target_class_ids, target_deltas, target_mask = \
            detection_target_layer(rpn_rois, gt_class_ids, gt_boxes, gt_masks, cfg)
        self.last_roi_counts = (int(p_rois.shape[0]), int(rois.shape[0]))
        empty = torch.zeros(0, device=dev)
        mrcnn_class_logits = mrcnn_bbox = mrcnn_mask = mrcnn_mask_logits = empty
        if rois.shape[0] > 0:
            mrcnn_class_logits, _, mrcnn_bbox = self.classifier(mrcnn_classifier_feature_maps, rois)
        if p_rois.shape[0] > 0:
            mrcnn_mask_logits, mrcnn_mask = self.mask(mrcnn_mask_feature_maps, p_rois)
        return [rpn_class_logits, rpn_bbox, target_class_ids, mrcnn_class_logits, target_deltas, mrcnn_bbox, target_mask,
                mrcnn_mask, mrcnn_mask_logits]

    def rpn_proposals(self, molded_images, mode, device_only=False):
        """backbone + FPN + RPN on both levels + proposal layer (reference model.py:1409-1437).  device_only: the last
        element is proposal_layer_device's (boxes, keep, count) instead of the [1,n,6] proposals (no read-back)."""
        cfg = self.config
        p2_out, p3_out = self.fpn(molded_images)
        layer_outputs = [self.rpn(p) for p in (p2_out, p3_out)]
        rpn_class_logits, rpn_class, rpn_bbox = [torch.cat(list(o), dim=1) for o in zip(*layer_outputs)]
        proposal_count = cfg.POST_NMS_ROIS_TRAINING if mode == "training" else cfg.POST_NMS_ROIS_INFERENCE
        layer = proposal_layer_device if device_only else proposal_layer
        rpn_rois = layer([rpn_class, rpn_bbox], proposal_count=proposal_count,
                         nms_threshold=cfg.RPN_NMS_THRESHOLD, anchors=self.anchors, config=cfg)
        return p2_out, p3_out, rpn_class_logits, rpn_class, rpn_bbox, rpn_rois

    def weighted_loss(self, losses):
        w = self.config.LOSS_WEIGHTS
        total = 0
        for name, l in zip(LOSS_NAMES, losses):
            total = total + w[name] * l
        return total

    def forward_backward(self, images, image_metas, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks):
        """One volume: predict('training') -> compute_losses -> weighted sum -> backward (reference model.py:1622-1640).
        Returns (total loss tensor, list of the six loss tensors); gradients accumulate into .grad.
        Same arithmetic as predict() + compute_losses(); the heads and their losses go through HeadsTail so that they can
        be replayed as CUDA graphs (enable_graphs)."""
        self.train()
        cfg = self.config
        dev = images.device
        frozen = bool(getattr(cfg, "STAGED_LOSSES", False)) and cfg.STAGE != 'beginning'    # LiTS: detector frozen, mask losses only
        # Everything up to the heads is enqueued without a host round trip; the single wait of the step is the read-back of
        # the proposal / positive / negative counts that size the two host-generator permutations (detection_targets_finish).
        with torch.set_grad_enabled(not frozen):
            p2, p3, rpn_class_logits, rpn_class, rpn_pred_bbox, prop = self.rpn_proposals(images, 'training', device_only=True)
        h, w, d = cfg.IMAGE_SHAPE[:3]
        if self._roi_scale is None or self._roi_scale.device != dev:
            self._roi_scale = _f32([d, h, w, d, h, w], dev)
        zero = _zero_loss(rpn_class_logits)
        rpn_counts = None
        cap = int(cfg.RPN_TRAIN_ANCHORS_PER_IMAGE)
        if frozen:
            rpn_class_loss = rpn_bbox_loss = zero
        else:
            rpn_class_loss, rpn_bbox_loss, rpn_counts = compute_rpn_losses_static(rpn_match, rpn_bbox, rpn_class_logits,
                                                                                  rpn_pred_bbox, cap)
        unet = self.mask.modified_u_net
        max_p = max(1, int(round(cfg.TRAIN_ROIS_PER_IMAGE * cfg.ROI_POSITIVE_RATIO)))
        drops_all = unet._drop_masks(max_p, dev)          # drawn for the largest possible positive count, sliced below
        state = detection_targets_begin(*prop, gt_boxes / self._roi_scale, cfg, extra_counts=rpn_counts)
        (p_rois, rois, target_class_ids, target_deltas, target_mask), extra = \
            detection_targets_finish(state, gt_class_ids, gt_masks, cfg)
        if rpn_counts is not None and (extra[0] > cap or extra[1] > min(cap, rpn_bbox.shape[1])):
            # more non-neutral anchors than build_rpn_targets ever emits: the fixed-size gather dropped some -- redo exactly
            rpn_class_loss = compute_rpn_class_loss(rpn_match, rpn_class_logits)
            rpn_bbox_loss = compute_rpn_bbox_loss(rpn_bbox, rpn_match, rpn_pred_bbox)
        P, R = int(p_rois.shape[0]), int(rois.shape[0])
        self.last_roi_counts = (P, R)
        if P > 0:
            if target_mask.dim() == 5:
                target_mask = torch.argmax(target_mask.long(), dim=1)
            if drops_all[0] is None:
                if self._unit_drops is None or self._unit_drops[0].device != dev or self._unit_drops[0].shape[0] < P:
                    self._unit_drops = [torch.ones((max(P, max_p), unet.base_n_filter * m), device=dev) for m in (1, 2, 4, 8, 16)]
                drops = [t[:P] for t in self._unit_drops]
            else:
                drops = [t[:P] for t in drops_all]
            if self._graphed_tails and ops.workspace_generation() != self._graph_ws_gen:
                # the shared conv workspace was reallocated after these graphs were captured (a larger RoI split or image):
                # they hold the freed buffer's address in kernel arguments and tensor maps -- drop them, re-capture lazily
                self._graphed_tails = {}
                self.graph_kernel_counts = {}
            self._tail.ensure_device(dev)
            tail = self._graphed_tails.get((P, R), self._tail) if self._graphed_tails is not None else self._tail
            if self._graphed_tails is not None and (P, R) not in self._graphed_tails:
                tail = self._capture_tail(P, R, p2, p3, images, rois, p_rois, target_class_ids, target_deltas, target_mask, drops)
            if self._graphed_tails is not None:
                self.graph_replays[(P, R)] = self.graph_replays.get((P, R), 0) + 1
            hl = tail(p2, p3, images, rois.detach(), p_rois.detach(), target_class_ids, target_deltas, target_mask, *drops)
            head_losses = [hl[0].reshape(1), hl[1].reshape(1), hl[2].reshape(1), hl[3].reshape(1)]
        elif R > 0:
            c_logits, _, c_bbox = self.classifier([p2, p3], rois)
            binary = (target_class_ids > 0).long()
            head_losses = [compute_mrcnn_class_loss(binary, c_logits), compute_mrcnn_bbox_loss(target_deltas, binary, c_bbox), zero, zero]
        else:
            head_losses = [zero, zero, zero, zero]
        losses = [rpn_class_loss, rpn_bbox_loss] + head_losses
        loss = self.weighted_loss(losses)
        if loss.requires_grad:
            loss.sum().backward()
        return loss, losses

    # -- CUDA graphs for the heads ----------------------------------------------------------------------------
    def enable_graphs(self, on=True):
        """Replay classifier + mask heads + their losses (forward and backward) as CUDA graphs, one pair per (P, R) RoI
        split, captured lazily on first use.  Call after the optimizer has re-homed the parameters (FlatSGD) and after one
        eager step (so that the conv workspace has reached its final size)."""
        self._graphed_tails = {} if on else None
        self._graph_ws_gen = ops.workspace_generation()
        self.graph_kernel_counts = {}
        self.graph_replays = {}

    def _capture_tail(self, P, R, *sample):
        from .ops import launch_count
        p2, p3, images, rois, p_rois, tcls, tdel, tmask, drops = sample
        args = (p2.detach().clone().requires_grad_(True), p3.detach().clone().requires_grad_(True), images.detach().clone(),
                rois.detach().clone(), p_rois.detach().clone(), tcls.clone(), tdel.clone(), tmask.clone()) + \
            tuple(dm.clone() for dm in drops)
        n0 = launch_count()
        # make_graphed_callables patches an nn.Module's forward in place: capture a fresh HeadsTail per RoI split (it only
        # references the shared classifier / mask modules) so that self._tail stays the eager path
        fresh = HeadsTail(self.classifier, self.mask, self.config.STAGE, self.config).ensure_device(p2.device)
        fresh.train(self.training)
        # Nothing may be destroyed while the capture is open: stale graphs dropped just before (workspace reallocation) sit in
        # reference cycles, and a cyclic collection that happened to run mid-capture would destroy them there (cudaGraphExec /
        # pool release are not capturable) -- collect now, keep the collector off until both graphs are captured.
        import gc
        gc.collect()
        torch.cuda.synchronize()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            graphed = torch.cuda.make_graphed_callables(fresh, args, num_warmup_iters=1, allow_unused_input=True)
        finally:
            if gc_was_on:
                gc.enable()
        # kernels per replay (forward + backward): launches during capture = (1 warm-up + 1 capture) iterations
        self.graph_kernel_counts[(P, R)] = (launch_count() - n0) // 2
        if self._graphed_tails and ops.workspace_generation() != self._graph_ws_gen:
            self._graphed_tails = {}          # the warm-up of this capture grew the workspace: older graphs are stale
        self._graphed_tails[(P, R)] = graphed
        self._graph_ws_gen = ops.workspace_generation()
        return graphed

    def train_step_device(self, optimizer, vol_i16, label_hwd, rpn_match, rpn_bbox, gt_boxes, gt_class_ids):
        """One optimizer step on one volume whose raw inputs are already on the device: mold (int16 -> normalised fp32,
        HWD -> DHW), forward, six losses, backward, global-norm clip + SGD.  Returns the 7 loss scalars as one tensor
        (total first) without synchronising."""
        image = ops.mold_volume_i16(vol_i16)
        label = label_hwd.permute(2, 0, 1).to(torch.int32).contiguous()
        optimizer.zero_grad()
        loss, losses = self.forward_backward(image, None, rpn_match.view(1, -1, 1), rpn_bbox.unsqueeze(0),
                                             gt_class_ids.unsqueeze(0), gt_boxes.unsqueeze(0), label)
        optimizer.step()
        return torch.stack([loss.detach().reshape(())] + [l.detach().reshape(()) for l in losses])

    def train_step_from_volume(self, optimizer, vol_i16, label_hwd, keys_pos=None, keys_neg=None):
        """train_step_device with the per-step targets generated ON THE DEVICE: GT box of the label
        volume (load_image_gt, reference model.py:1058-1076) and RPN match / delta targets (build_rpn_targets,
        model.py:1090-1181).  Inputs: the raw int16 scan [H,W,D] and its class-id label [H,W,D], both on the device."""
        from . import targets
        cfg = self.config
        label_dhw = label_hwd.permute(2, 0, 1).to(torch.int32).contiguous()
        gt_boxes = targets.gt_box_from_label(label_dhw, cfg.NUM_CLASSES)
        rpn_match, rpn_bbox = targets.build_rpn_targets(self.anchors, gt_boxes[:1], cfg, keys_pos, keys_neg)
        gt_class_ids = torch.arange(1, cfg.NUM_CLASSES, dtype=torch.int32, device=vol_i16.device)
        return self.train_step_device(optimizer, vol_i16, label_hwd, rpn_match, rpn_bbox, gt_boxes, gt_class_ids)

    def train_step_from_host(self, optimizer, step_inputs, wait=True):
        """The end-to-end call: pinned host buffers of one volume -> H2D -> train_step_device -> D2H of the 7 losses.
        wait=False returns a PendingLosses instead of blocking: the copy into pinned host memory is queued behind the step
        and .result() waits for it -- a training loop reads step i's losses while step i+1 is already running, so the host
        never leaves the device without queued work between steps."""
        dev = self.anchors.device
        args = [t.to(dev, non_blocking=True) for t in step_inputs.tensors()]
        out = self.train_step_device(optimizer, *args)
        pending = PendingLosses(out)
        return pending.result() if wait else pending

    # -- inference ------------------------------------------------------------------------------------------
    def detect(self, images):
        """reference model.py:1341-1389.  images: list of [H,W,D,C] arrays (the raw scans).  The resize + normalisation
        before the network and the mask resize + paste + argmax after it run on the device; what
        crosses PCIe is the raw scan in, the detections [n,8] and one uint8 class-id volume out."""
        start_time = time.time()
        molded, image_metas, windows = self.mold_inputs_device(images)
        with torch.no_grad():
            detections, mrcnn_mask = self.predict([molded, image_metas], mode='inference')
        detections = detections.detach().cpu().numpy()
        print("detect done, using time", time.time() - start_time)
        results = []
        for i, image in enumerate(images):
            rois, class_ids, scores, mask = self.unmold_detections(
                detections[i], mrcnn_mask[i], [image.shape[3], image.shape[2], image.shape[0], image.shape[1]], windows[i])
            results.append({"rois": rois, "class_ids": class_ids, "scores": scores, "mask": mask})
        return results

    def mold_inputs_device(self, images):
        """mold_inputs (reference model.py:1774-1810) with the arithmetic on the device: order-1 resize to
        [IMAGE_MAX_DIM, IMAGE_MAX_DIM, IMAGE_MIN_DIM] ('self' mode), cast back to the scan dtype, (x - mean) / std,
        [H,W,D,C] -> [C,D,H,W].  Returns (molded device tensor [N,1,D,H,W], image_metas, windows)."""
        c = self.config
        dev = self.anchors.device
        molded, metas, windows = [], [], []
        for image in images:
            if image.shape[3] != 1:
                raise RuntimeError("single-channel scans only (IMAGE_CHANNEL_COUNT = 1 in every CFUN config)")
            h, w, d = image.shape[:3]
            if c.IMAGE_RESIZE_MODE == "none":
                tgt, window = (h, w, d), (0, 0, 0, d, h, w)
            elif c.IMAGE_RESIZE_MODE == "self":
                tgt, window = (c.IMAGE_MAX_DIM, c.IMAGE_MAX_DIM, c.IMAGE_MIN_DIM), (0, 0, 0, c.IMAGE_MIN_DIM, c.IMAGE_MAX_DIM, c.IMAGE_MAX_DIM)
            else:
                raise NotImplementedError("IMAGE_RESIZE_MODE %r is not used by the CFUN configs" % c.IMAGE_RESIZE_MODE)
            vol = np.ascontiguousarray(image[..., 0])
            if vol.dtype == np.int16:
                v = torch.from_numpy(vol).to(dev)
                if tuple(tgt) != (h, w, d):
                    v = ops.resize_linear3d(v, tgt)
                m = ops.mold_volume_i16(v)                                      # [1,1,D,H,W]
            else:
                v = torch.from_numpy(vol.astype(np.float32)).to(dev)
                if tuple(tgt) != (h, w, d):
                    v = ops.resize_linear3d(v, tgt)
                if np.issubdtype(vol.dtype, np.integer):
                    v = torch.trunc(v)                                          # .astype(image_dtype) of the reference
                v = (v - v.mean()) / v.std(unbiased=False)
                m = v.permute(2, 0, 1)[None, None].contiguous()
            molded.append(m)
            windows.append(window)
            metas.append(compose_image_meta(0, image.shape, window, np.zeros([c.NUM_CLASSES], dtype=np.int32)))
        return torch.cat(molded, 0), np.stack(metas), np.stack(windows)

    def mold_inputs(self, images):
        """reference signature (numpy out): the device computation of mold_inputs_device, copied back"""
        molded, metas, windows = self.mold_inputs_device(images)
        return molded.cpu().numpy(), metas, windows

    def unmold_detections(self, detections, mrcnn_mask, image_shape, window):
        """reference model.py:1812-1864.  detections [n,8] numpy; mrcnn_mask the class probabilities of every detection,
        either the device tensor predict() returned ([n,ncls,d,h,w]) or the reference's numpy layout [n,d,h,w,ncls];
        image_shape [C,D,H,W] of the original scan.  The box arithmetic is a handful of integers (host); the mask resize +
        paste + argmax is one device kernel (ops.unmold_mask_argmax)."""
        zero_ix = np.where(detections[:, 6] == 0)[0]
        N = zero_ix[0] if zero_ix.shape[0] > 0 else detections.shape[0]
        boxes = detections[:N, :6].astype(np.int32)
        scores = detections[:N, 7]
        keep = np.arange(N)
        sc = np.array([image_shape[1] / (window[3] - window[0]), image_shape[2] / (window[4] - window[1]),
                       image_shape[3] / (window[5] - window[2])] * 2)
        sh = np.array(list(window[:3]) * 2)
        boxes = np.multiply(boxes - sh, sc).astype(np.int32)
        bad = np.where((boxes[:, 3] - boxes[:, 0]) * (boxes[:, 4] - boxes[:, 1]) * (boxes[:, 5] - boxes[:, 2]) <= 0)[0]
        if bad.shape[0] > 0:
            boxes, scores, keep = np.delete(boxes, bad, 0), np.delete(scores, bad, 0), np.delete(keep, bad, 0)
        if torch.is_tensor(mrcnn_mask):
            m0 = mrcnn_mask[int(keep[0])]                                        # [ncls,d,h,w]
        else:
            m0 = torch.from_numpy(np.ascontiguousarray(mrcnn_mask[int(keep[0])])).to(self.anchors.device).permute(3, 0, 1, 2)
        full_mask = ops.unmold_mask_argmax(m0, boxes[0], image_shape[1:4]).cpu().numpy().astype(np.int64)
        boxes[:, [0, 1, 2, 3, 4, 5]] = boxes[:, [1, 2, 0, 4, 5, 3]]
        return boxes, np.arange(1, 8), scores, full_mask

    # -- training loop ------------------------------------------------------------------------------------------
    def make_optimizer(self, learning_rate):
        """SGD(momentum) with weight decay on every trainable parameter whose name lacks 'bn' (reference :1538-1545),
        as a fused clip + step kernel over flat buffers (cfun_b200.dp.FlatSGD)."""
        from .dp import FlatSGD
        return FlatSGD(self, lr=learning_rate, momentum=self.config.LEARNING_MOMENTUM,
                       weight_decay=self.config.WEIGHT_DECAY, clip_norm=5.0)

    def train_model(self, train_dataset, val_dataset, learning_rate, epochs):
        """reference model.py:1516-1572"""
        train_gen = torch.utils.data.DataLoader(Dataset(train_dataset, self.config), batch_size=1, shuffle=True,
                                                num_workers=getattr(self.config, "LOADER_WORKERS", 4))
        val_gen = torch.utils.data.DataLoader(Dataset(val_dataset, self.config), batch_size=1, shuffle=True,
                                              num_workers=getattr(self.config, "LOADER_WORKERS", 4))
        self.set_trainable(".*")
        optimizer = self.make_optimizer(learning_rate)
        start_datetime = time.strftime("%Y-%m-%d %H:%M:%S", time.localtime())
        out_dir = os.path.join("./logs/heart", start_datetime)
        os.makedirs(out_dir, exist_ok=True)
        total_start = time.time()
        for epoch in range(self.epoch + 1, epochs + 1):
            log("Epoch {}/{}.".format(epoch, epochs))
            t0 = time.time()
            angle = np.random.randint(-20, 21)
            stats = self.train_epoch(train_gen, optimizer, self.config.STEPS_PER_EPOCH, angle, train_dataset)
            print("One Training Epoch time:", int(time.time() - t0), "Total time:", int(time.time() - total_start))
            if epoch % 5 == 0:
                vstats = self.valid_epoch(val_gen, self.config.VALIDATION_STEPS, angle, val_dataset)
                torch.save(self.state_dict(), os.path.join(out_dir, "model%d_loss: %s_val: %s" % (
                    epoch, round(stats[0], 4), round(vstats[0], 4))))
        self.epoch = epochs

    def _batch_to_device(self, inputs, angle, dataset):
        image = inputs[0].squeeze(0).cpu().numpy()
        mask = inputs[2].squeeze(0).cpu().numpy()
        images, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks = load_image_gt(
            image, mask, angle, dataset, self.config, self.anchors.cpu().numpy())
        dev = self.anchors.device
        return (torch.from_numpy(images).float().unsqueeze(0).to(dev), inputs[1].numpy(),
                torch.from_numpy(rpn_match).unsqueeze(0).to(dev), torch.from_numpy(rpn_bbox).float().unsqueeze(0).to(dev),
                torch.from_numpy(gt_class_ids).unsqueeze(0).to(dev), torch.from_numpy(gt_boxes).float().unsqueeze(0).to(dev),
                torch.from_numpy(gt_masks).float().unsqueeze(0).to(dev))

    def train_epoch(self, datagenerator, optimizer, steps, angle, dataset):
        sums = np.zeros(7)
        batch_count, step = 0, 0
        optimizer.zero_grad()
        for inputs in datagenerator:
            batch_count += 1
            images, metas, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks = self._batch_to_device(inputs, angle, dataset)
            loss, losses = self.forward_backward(images, metas, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks)
            # reference: clip_grad_norm_ on the ACCUMULATED gradient after every backward, optimizer step every BATCH_SIZE
            # volumes (model.py:1641-1645).  The clip of the stepping iteration is fused into FlatSGD.step().
            if (batch_count % self.config.BATCH_SIZE) == 0:
                optimizer.step()
                optimizer.zero_grad()
                batch_count = 0
            else:
                optimizer.clip_()
            vals = torch.stack([loss.detach().reshape(())] + [l.detach().reshape(()) for l in losses]).cpu().numpy()
            sums += vals / steps
            if step == steps - 1:
                break
            step += 1
        return tuple(sums)

    def valid_epoch(self, datagenerator, steps, angle, dataset):
        sums = np.zeros(7)
        step = 0
        for inputs in datagenerator:
            images, metas, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks = self._batch_to_device(inputs, angle, dataset)
            with torch.no_grad():
                outs = self.predict([images, metas, gt_class_ids, gt_boxes, gt_masks], mode='training')
                if outs[2].shape[0] == 0:
                    continue
                losses = compute_losses(rpn_match, rpn_bbox, outs[0], outs[1], outs[2], outs[3], outs[4], outs[5], outs[6],
                                        outs[7], outs[8], self.config.STAGE)
                loss = self.weighted_loss(losses)
            sums += torch.stack([loss.reshape(())] + [l.reshape(()) for l in losses]).cpu().numpy() / steps
            if step == steps - 1:
                break
            step += 1
        return tuple(sums)
