"""CPU oracle for the CFUN volumetric hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
path (cfun_b200/) never imports anything from oracle/ and fails loudly when its
CUDA library is missing.

It is a CPU restatement (numpy for integer/box arithmetic, torch-CPU fp32 for the
floating-point layers) of the reference algorithm, each function citing the
reference file:line it follows.  Pinning: tests/test_oracle_golden.py checks every
function here against golden vectors produced by the *unmodified* reference run in
the build container (oracle/gen_golden.py -> tests/golden/*.npz); when
/root/reference is present the same tests also run the live reference beside it.
One boundary is "parity unpinned": skimage.transform.resize(order=0) used for the
mask targets (reference model.py:490) -- scikit-image is absent and unpinned in the
reference (README.md:16); `nn_resize` restates the half-pixel-centre nearest
neighbour map of scipy.ndimage.zoom(order=0, mode='grid-constant', grid_mode=True) that skimage>=0.19
delegates to, and is pinned against scipy here.

Deliberate, documented deviations from undefined reference behaviour:
  * sort ties: the reference sorts with numpy argsort()[::-1] / torch.sort, whose
    tie order is unspecified.  The oracle (and the CUDA path) define the total
    order (score descending, index ascending).
  * RNG: torch.randperm / Dropout3d draws are *inputs* here (perm / mask arguments)
    so CPU and CUDA paths can be compared on identical draws.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# anchors  (reference utils.py:467-528, model.py:91-101)
# --------------------------------------------------------------------------------------


def backbone_shapes(image_shape, strides):
    """model.py:91-101: image_shape = (H, W, D, C); rows are (depth, height, width)."""
    H, W, D = image_shape[:3]
    return np.array([[int(math.ceil(D / s)), int(math.ceil(H / s)), int(math.ceil(W / s))] for s in strides])


def generate_anchors(scale, shape, feature_stride, anchor_stride=1):
    """utils.py:467-508 with one ratio.  np.meshgrid's default 'xy' indexing makes the
    flat enumeration y-slowest, z-middle, x-fastest."""
    sz = np.arange(0, shape[0], anchor_stride) * feature_stride
    sy = np.arange(0, shape[1], anchor_stride) * feature_stride
    sx = np.arange(0, shape[2], anchor_stride) * feature_stride
    out = np.empty((len(sy), len(sz), len(sx), 6), dtype=np.float64)
    half = 0.5 * float(scale)
    out[..., 0] = sz[None, :, None] - half
    out[..., 1] = sy[:, None, None] - half
    out[..., 2] = sx[None, None, :] - half
    out[..., 3] = sz[None, :, None] + half
    out[..., 4] = sy[:, None, None] + half
    out[..., 5] = sx[None, None, :] + half
    return out.reshape(-1, 6)


def generate_pyramid_anchors(scales, feature_shapes, feature_strides, anchor_stride=1):
    """utils.py:511-528."""
    return np.concatenate([generate_anchors(scales[i], feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


# --------------------------------------------------------------------------------------
# IoU / NMS  (reference utils.py:50-70, 122-157) -- fp32, unfused, bit-exact contract
# --------------------------------------------------------------------------------------


def box_volume(boxes):
    """utils.py:137: (z2-z1)*(y2-y1)*(x2-x1), left to right, fp32."""
    b = np.asarray(boxes, dtype=np.float32)
    return ((b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1])) * (b[:, 5] - b[:, 2])


def compute_iou(box, boxes, box_vol, boxes_vol):
    """utils.py:50-70.  All arithmetic fp32; 1e-6 and 0 are weak scalars -> fp32."""
    z1 = np.maximum(box[0], boxes[:, 0])
    z2 = np.minimum(box[3], boxes[:, 3])
    y1 = np.maximum(box[1], boxes[:, 1])
    y2 = np.minimum(box[4], boxes[:, 4])
    x1 = np.maximum(box[2], boxes[:, 2])
    x2 = np.minimum(box[5], boxes[:, 5])
    zero = np.float32(0)
    inter = (np.maximum(x2 - x1, zero) * np.maximum(y2 - y1, zero)) * np.maximum(z2 - z1, zero)
    union = (box_vol + boxes_vol) - inter
    return inter / (union + np.float32(1e-6))


def sort_desc(scores):
    """Total order used everywhere on this path: score descending, index ascending."""
    scores = np.asarray(scores, dtype=np.float32)
    return np.argsort(-scores, kind="stable")


def non_max_suppression(boxes, scores, threshold, max_num):
    """utils.py:122-157: greedy; stop as soon as len(pick) >= max_num; drop iou > thr (strict)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    vol = box_volume(boxes)
    thr = np.float32(threshold)
    ixs = sort_desc(scores)
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        if len(pick) >= max_num:
            break
        rest = ixs[1:]
        iou = compute_iou(boxes[i], boxes[rest], vol[i], vol[rest])
        ixs = rest[~(iou > thr)]
    return np.array(pick, dtype=np.int32)


def compute_overlaps(boxes1, boxes2):
    """utils.py:73-89 (float64 in the reference because anchors are float64 there)."""
    boxes1 = np.asarray(boxes1)
    boxes2 = np.asarray(boxes2)
    v1 = (boxes1[:, 3] - boxes1[:, 0]) * (boxes1[:, 4] - boxes1[:, 1]) * (boxes1[:, 5] - boxes1[:, 2])
    v2 = (boxes2[:, 3] - boxes2[:, 0]) * (boxes2[:, 4] - boxes2[:, 1]) * (boxes2[:, 5] - boxes2[:, 2])
    out = np.zeros((boxes1.shape[0], boxes2.shape[0]))
    for j in range(boxes2.shape[0]):
        b = boxes2[j]
        z1 = np.maximum(b[0], boxes1[:, 0]); z2 = np.minimum(b[3], boxes1[:, 3])
        y1 = np.maximum(b[1], boxes1[:, 1]); y2 = np.minimum(b[4], boxes1[:, 4])
        x1 = np.maximum(b[2], boxes1[:, 2]); x2 = np.minimum(b[5], boxes1[:, 5])
        inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0) * np.maximum(z2 - z1, 0)
        out[:, j] = inter / (v2[j] + v1 - inter + 1e-6)
    return out


# --------------------------------------------------------------------------------------
# box decode / clip / proposals  (reference model.py:155-258)
# --------------------------------------------------------------------------------------


def apply_box_deltas(boxes, deltas):
    """model.py:155-182, same op order, fp32 torch CPU."""
    d = boxes[:, 3] - boxes[:, 0]
    h = boxes[:, 4] - boxes[:, 1]
    w = boxes[:, 5] - boxes[:, 2]
    cz = boxes[:, 0] + 0.5 * d
    cy = boxes[:, 1] + 0.5 * h
    cx = boxes[:, 2] + 0.5 * w
    cz = cz + deltas[:, 0] * d
    cy = cy + deltas[:, 1] * h
    cx = cx + deltas[:, 2] * w
    d = d * torch.exp(deltas[:, 3])
    h = h * torch.exp(deltas[:, 4])
    w = w * torch.exp(deltas[:, 5])
    z1 = cz - 0.5 * d
    y1 = cy - 0.5 * h
    x1 = cx - 0.5 * w
    return torch.stack([z1, y1, x1, z1 + d, y1 + h, x1 + w], dim=1)


def clip_boxes(boxes, window):
    """model.py:185-196."""
    lo = torch.tensor([window[0], window[1], window[2]] * 2, dtype=boxes.dtype)
    hi = torch.tensor([window[3], window[4], window[5]] * 2, dtype=boxes.dtype)
    return torch.max(torch.min(boxes, hi), lo)


def proposal_layer(rpn_probs, rpn_bbox, anchors, proposal_count, nms_threshold, pre_nms_limit, image_shape,
                   std_dev=(0.1, 0.1, 0.1, 0.2, 0.2, 0.2)):
    """model.py:199-258.  rpn_probs [A,2], rpn_bbox [A,6], anchors [A,6] fp32; image_shape (H,W,D,..).
    Returns (normalized boxes [n,6], kept indices into the pre-NMS top-k list, top-k order)."""
    scores = rpn_probs[:, 1]
    deltas = rpn_bbox * torch.tensor(std_dev, dtype=torch.float32).view(1, 6)
    k = min(pre_nms_limit, anchors.shape[0])
    order = torch.from_numpy(sort_desc(scores.detach().numpy())[:k].copy())
    top_scores = scores[order]
    boxes = apply_box_deltas(anchors[order], deltas[order])
    H, W, D = [int(v) for v in image_shape[:3]]
    boxes = clip_boxes(boxes, [0.0, 0.0, 0.0, float(D), float(H), float(W)])
    keep = non_max_suppression(boxes.detach().numpy(), top_scores.detach().numpy(), nms_threshold, proposal_count)
    keep_t = torch.from_numpy(keep.astype(np.int64))
    norm = torch.tensor([D, H, W, D, H, W], dtype=torch.float32)
    return boxes[keep_t] / norm, keep, order.numpy()


# --------------------------------------------------------------------------------------
# RoI crop-resize  (reference model.py:265-370)
# --------------------------------------------------------------------------------------


def roi_align(feature_map, pool_size, boxes):
    """model.py:265-289: denorm by the feature-map size, floor lower / ceil upper, python-slice crop,
    trilinear align_corners=True resize; any failure (empty crop) leaves zeros.
    feature_map [C,D,H,W]; boxes [n,6] normalised."""
    C, D, H, W = feature_map.shape
    scale = torch.tensor([D, H, W, D, H, W], dtype=torch.float32)
    b = boxes.detach() * scale
    b = torch.cat([b[:, :3].floor(), b[:, 3:].ceil()], dim=1).long()
    out = torch.zeros((b.shape[0], C) + tuple(pool_size), dtype=feature_map.dtype)
    rows = []
    for i in range(b.shape[0]):
        z1, y1, x1, z2, y2, x2 = [int(v) for v in b[i]]
        crop = feature_map[:, z1:z2, y1:y2, x1:x2]
        if crop.numel() == 0:
            rows.append(out[i])
            continue
        rows.append(F.interpolate(crop.unsqueeze(0), size=tuple(pool_size), mode="trilinear", align_corners=True)[0])
    return torch.stack(rows, 0) if rows else out


def roi_level(boxes):
    """model.py:322-332: level = clamp(round(4 + log2(h*w*d)/3), 2, 3) on normalised boxes, fp32,
    log2 = log(x)/log(2), round-half-even."""
    d = boxes[:, 3] - boxes[:, 0]
    h = boxes[:, 4] - boxes[:, 1]
    w = boxes[:, 5] - boxes[:, 2]
    ln2 = torch.log(torch.tensor([2.0], dtype=torch.float32))
    lvl = 4 + (1.0 / 3.0) * (torch.log(h * w * d) / ln2)
    return lvl.round().int().clamp(2, 3)


def pyramid_roi_align(boxes, feature_maps, pool_size):
    """model.py:292-370 for batch 1.  boxes [n,6] normalised; feature_maps list of [C,D,H,W] (P2, P3).
    Output rows are in the original box order."""
    lvl = roi_level(boxes.detach())
    out = None
    for i, level in enumerate((2, 3)):
        ix = torch.nonzero(lvl == level)[:, 0]
        if ix.numel() == 0:
            continue
        pooled = roi_align(feature_maps[i], pool_size, boxes[ix].detach())
        if out is None:
            out = torch.zeros((boxes.shape[0],) + tuple(pooled.shape[1:]), dtype=pooled.dtype)
        out = out.index_copy(0, ix, pooled)
    return out


# --------------------------------------------------------------------------------------
# detection targets  (reference model.py:377-563, utils.py:92-119, 318-339)
# --------------------------------------------------------------------------------------


def bbox_overlaps(boxes1, boxes2):
    """model.py:377-411: all pairs, no epsilon, fp32."""
    b1 = boxes1[:, None, :]
    b2 = boxes2[None, :, :]
    z1 = torch.max(b1[..., 0], b2[..., 0]); y1 = torch.max(b1[..., 1], b2[..., 1]); x1 = torch.max(b1[..., 2], b2[..., 2])
    z2 = torch.min(b1[..., 3], b2[..., 3]); y2 = torch.min(b1[..., 4], b2[..., 4]); x2 = torch.min(b1[..., 5], b2[..., 5])
    zero = torch.zeros((), dtype=boxes1.dtype)
    inter = torch.max(x2 - x1, zero) * torch.max(y2 - y1, zero) * torch.max(z2 - z1, zero)
    v1 = (b1[..., 3] - b1[..., 0]) * (b1[..., 4] - b1[..., 1]) * (b1[..., 5] - b1[..., 2])
    v2 = (b2[..., 3] - b2[..., 0]) * (b2[..., 4] - b2[..., 1]) * (b2[..., 5] - b2[..., 2])
    return inter / (v1 + v2 - inter)


def box_refinement(box, gt_box):
    """utils.py:92-119."""
    d = box[:, 3] - box[:, 0]; h = box[:, 4] - box[:, 1]; w = box[:, 5] - box[:, 2]
    cz = box[:, 0] + 0.5 * d; cy = box[:, 1] + 0.5 * h; cx = box[:, 2] + 0.5 * w
    gd = gt_box[:, 3] - gt_box[:, 0]; gh = gt_box[:, 4] - gt_box[:, 1]; gw = gt_box[:, 5] - gt_box[:, 2]
    gz = gt_box[:, 0] + 0.5 * gd; gy = gt_box[:, 1] + 0.5 * gh; gx = gt_box[:, 2] + 0.5 * gw
    return torch.stack([(gz - cz) / d, (gy - cy) / h, (gx - cx) / w,
                        torch.log(gd / d), torch.log(gh / h), torch.log(gw / w)], dim=1)


def nn_index(out_size, in_size):
    """Half-pixel-centre nearest neighbour source index for every output index:
    floor((o + 0.5) * in / out), the order-0 map of scipy.ndimage.zoom(grid_mode=True)
    that skimage>=0.19 resize(order=0, anti_aliasing=False) delegates to."""
    o = np.arange(out_size, dtype=np.float64)
    idx = np.floor((o + 0.5) * (float(in_size) / float(out_size)) - 0.5 + 0.5).astype(np.int64)
    return np.clip(idx, 0, in_size - 1)


def nn_resize(arr, out_shape):
    """utils.py:318-339 with order=0 (see module docstring: parity unpinned boundary)."""
    arr = np.asarray(arr)
    assert arr.ndim == len(out_shape)
    res = arr
    for ax, (o, i) in enumerate(zip(out_shape, arr.shape)):
        res = np.take(res, nn_index(o, i), axis=ax)
    return res


def mask_crop_window(roi, mask_dhw):
    """model.py:483-488: int() truncation of dim * normalised coord (fp32 product)."""
    D, H, W = mask_dhw
    r = roi.detach().to(torch.float32)
    z1 = int(D * r[0]); z2 = int(D * r[3])
    y1 = int(H * r[1]); y2 = int(H * r[4])
    x1 = int(W * r[2]); x2 = int(W * r[5])
    return z1, y1, x1, z2, y2, x2


def detection_target_layer(proposals, gt_class_ids, gt_boxes, gt_masks, mask_shape, rois_per_image, positive_ratio,
                           iou_threshold, bbox_std_dev, perm_pos=None, perm_neg=None):
    """model.py:414-563 for batch 1.  proposals [N,6] normalised, gt_boxes [G,6] normalised,
    gt_masks [8,D,H,W].  perm_pos / perm_neg stand in for the two torch.randperm draws
    (model.py:459,505): 1-D int64 permutations of the candidate counts.
    Returns (positive_rois, rois, class_ids, deltas, masks[float64])."""
    overlaps = bbox_overlaps(proposals, gt_boxes)
    iou_max = overlaps.max(dim=1)[0]
    pos_ix = torch.nonzero(iou_max >= iou_threshold)[:, 0]
    neg_ix = torch.nonzero(iou_max < iou_threshold)[:, 0]
    positive_count = 0
    empty = torch.zeros((0,), dtype=torch.float32)
    if pos_ix.numel() > 0:
        want = int(rois_per_image * positive_ratio)
        perm = perm_pos if perm_pos is not None else torch.randperm(pos_ix.numel())
        assert perm.numel() == pos_ix.numel()
        pos_ix = pos_ix[perm[:want]]
        positive_count = pos_ix.numel()
        positive_rois = proposals[pos_ix]
        assign = overlaps[pos_ix].max(dim=1)[1]
        roi_gt_boxes = gt_boxes[assign]
        roi_gt_class_ids = gt_class_ids[assign]
        deltas = box_refinement(positive_rois.detach(), roi_gt_boxes.detach())
        deltas = deltas / torch.tensor(bbox_std_dev, dtype=torch.float32)
        gm = gt_masks.detach().numpy()
        masks = np.zeros((positive_count, gm.shape[0]) + tuple(mask_shape))
        for i in range(positive_count):
            z1, y1, x1, z2, y2, x2 = mask_crop_window(positive_rois[i], gm.shape[1:])
            crop = gm[:, z1:z2, y1:y2, x1:x2]
            masks[i] = nn_resize(crop, (gm.shape[0],) + tuple(mask_shape))
        masks = torch.from_numpy(masks).double()
    negative_count = 0
    if neg_ix.numel() > 0 and positive_count > 0:
        want = int((1.0 / positive_ratio) * positive_count - positive_count)
        perm = perm_neg if perm_neg is not None else torch.randperm(neg_ix.numel())
        assert perm.numel() == neg_ix.numel()
        neg_ix = neg_ix[perm[:want]]
        negative_count = neg_ix.numel()
        negative_rois = proposals[neg_ix]
    if positive_count > 0 and negative_count > 0:
        rois = torch.cat([positive_rois, negative_rois], 0)
        class_ids = torch.cat([roi_gt_class_ids.long(), torch.zeros(negative_count, dtype=torch.long)], 0)
        deltas = torch.cat([deltas, torch.zeros(negative_count, 6)], 0)
        return positive_rois, rois, class_ids, deltas, masks
    if positive_count > 0:
        return positive_rois, positive_rois, roi_gt_class_ids.long(), deltas, masks
    # no positives: model.py:534-562 (negatives need positive_count > 0, so everything is empty)
    return empty.view(0, 6), empty.view(0, 6), torch.zeros(0, dtype=torch.long), empty.view(0, 6), empty


# --------------------------------------------------------------------------------------
# detection refinement (inference)  (reference model.py:570-693)
# --------------------------------------------------------------------------------------


def refine_detections(rois, probs, deltas, window, image_shape, min_confidence, nms_threshold, max_instances,
                      std_dev=(0.1, 0.1, 0.1, 0.2, 0.2, 0.2)):
    """model.py:584-676.  Returns [n,8] (z1,y1,x1,z2,y2,x2,class,score); raises like the reference
    (UnboundLocalError there) when nothing survives -- callers must handle it."""
    class_ids = probs.argmax(dim=1)
    idx = torch.arange(class_ids.shape[0])
    class_scores = probs[idx, class_ids]
    deltas_specific = deltas[idx, class_ids]
    refined = apply_box_deltas(rois, deltas_specific * torch.tensor(std_dev, dtype=torch.float32).view(1, 6))
    H, W, D = [int(v) for v in image_shape[:3]]
    refined = refined * torch.tensor([D, H, W, D, H, W], dtype=torch.float32)
    refined = clip_boxes(refined, [float(w) for w in window])
    refined = torch.round(refined)
    keep_bool = class_ids > 0
    if min_confidence:
        keep_bool = keep_bool & (class_scores >= min_confidence)
    keep = torch.nonzero(keep_bool)[:, 0]
    if keep.numel() == 0:
        raise RuntimeError("no detection survives (reference raises UnboundLocalError at model.py:662)")
    nms_keep = []
    pre_cls = class_ids[keep]; pre_scores = class_scores[keep]; pre_rois = refined[keep]
    for cid in torch.unique(pre_cls):
        ixs = torch.nonzero(pre_cls == cid)[:, 0]
        order = torch.from_numpy(sort_desc(pre_scores[ixs].detach().numpy()).copy())
        ck = non_max_suppression(pre_rois[ixs][order].detach().numpy(), pre_scores[ixs][order].detach().numpy(),
                                 nms_threshold, max_instances)
        nms_keep.append(keep[ixs[order[torch.from_numpy(ck.astype(np.int64))]]])
    nms_keep = torch.unique(torch.cat(nms_keep))
    keep = torch.from_numpy(np.intersect1d(keep.numpy(), nms_keep.numpy()))
    k = min(max_instances, keep.numel())
    top = torch.from_numpy(sort_desc(class_scores[keep].detach().numpy())[:k].copy())
    keep = keep[top]
    return torch.cat([refined[keep], class_ids[keep].unsqueeze(1).float(), class_scores[keep].unsqueeze(1)], dim=1)


# --------------------------------------------------------------------------------------
# network layers, functional over a reference-keyed state_dict (fp32, NCDHW)
# --------------------------------------------------------------------------------------


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv3d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _bn_eval(sd, name, x, eps=1e-5):
    """BatchNorm3d is frozen in eval mode on this path (model.py:1297-1304, 1401-1406)."""
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, eps)


def bottleneck(sd, p, x, block, expand, stride):
    """backbone.py:26-114."""
    st = "ABC"[(block - 1) % 3]
    out = F.relu(_bn_eval(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride=stride)))
    if st == "A":
        out = F.relu(_bn_eval(sd, p + ".bn2", _conv(sd, p + ".conv2", out, padding=(0, 1, 1))))
        out = F.relu(_bn_eval(sd, p + ".bn3", _conv(sd, p + ".conv3", out, padding=(1, 0, 0))))
    elif st == "B":
        s = F.relu(_bn_eval(sd, p + ".bn2", _conv(sd, p + ".conv2", out, padding=(0, 1, 1))))
        t = F.relu(_bn_eval(sd, p + ".bn3", _conv(sd, p + ".conv3", out, padding=(1, 0, 0))))
        out = t + s
    else:
        s = F.relu(_bn_eval(sd, p + ".bn2", _conv(sd, p + ".conv2", out, padding=(0, 1, 1))))
        t = F.relu(_bn_eval(sd, p + ".bn3", _conv(sd, p + ".conv3", s, padding=(1, 0, 0))))
        out = s + t
    out = _bn_eval(sd, p + ".bn4", _conv(sd, p + ".conv4", out))
    res = x
    if expand:
        res = _bn_eval(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=2))
    return F.relu(out + res)


def fpn_forward(sd, x, layers=(2, 3), stem_pad=(1, 3, 3), prefix="fpn"):
    """backbone.py:117-158 (C1..C3) + model.py:136-148 (FPN)."""
    x = _conv(sd, prefix + ".C1.0", x, stride=2, padding=stem_pad)
    x = F.max_pool3d(F.relu(_bn_eval(sd, prefix + ".C1.1", x)), 2, 2)
    for b in range(layers[0]):
        x = bottleneck(sd, "%s.C2.%d" % (prefix, b), x, b + 1, b == 0, 2 if b == 0 else 1)
    c2 = x
    for b in range(layers[1]):
        x = bottleneck(sd, "%s.C3.%d" % (prefix, b), x, b + 1, b == 0, 2 if b == 0 else 1)
    c3 = x
    p3 = _conv(sd, prefix + ".P3_conv1", c3)
    p2 = _conv(sd, prefix + ".P2_conv1", c2) + F.interpolate(p3, scale_factor=2, mode="nearest")
    p3 = _conv(sd, prefix + ".P3_conv2", p3, padding=1)
    p2 = _conv(sd, prefix + ".P2_conv2", p2, padding=1)
    return p2, p3


def rpn_forward(sd, x, prefix="rpn"):
    """model.py:719-743."""
    x = F.relu(_conv(sd, prefix + ".conv_shared", x, padding=1))
    logits = _conv(sd, prefix + ".conv_class", x).permute(0, 2, 3, 4, 1).contiguous().view(x.shape[0], -1, 2)
    probs = F.softmax(logits, dim=2)
    bbox = _conv(sd, prefix + ".conv_bbox", x).permute(0, 2, 3, 4, 1).contiguous().view(x.shape[0], -1, 6)
    return logits, probs, bbox


def classifier_forward(sd, pooled, prefix="classifier"):
    """model.py:768-784 after the RoI crop."""
    x = F.relu(_bn_eval(sd, prefix + ".bn1", _conv(sd, prefix + ".conv1", pooled), eps=1e-3))
    x = F.relu(_bn_eval(sd, prefix + ".bn2", _conv(sd, prefix + ".conv2", x), eps=1e-3))
    x = x.view(-1, x.shape[1])
    logits = F.linear(x, sd[prefix + ".linear_class.weight"], sd[prefix + ".linear_class.bias"])
    probs = F.softmax(logits, dim=1)
    bbox = F.linear(x, sd[prefix + ".linear_bbox.weight"], sd[prefix + ".linear_bbox.bias"])
    return logits, probs, bbox.view(bbox.shape[0], -1, 6)


def _inorm(x):
    return F.instance_norm(x, eps=1e-5)


def _lrelu(x):
    return F.leaky_relu(x, 0.01)


def _up(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def unet_forward(sd, x, stage="beginning", drop=None, prefix="mask.modified_u_net"):
    """mask_branch.py:124-220.  drop: None (eval) or list of 5 per-call channel masks [N,C,1,1,1]
    already scaled by 1/(1-p) (Dropout3d p=0.6, mask_branch.py:19)."""
    w = lambda n: sd[prefix + "." + n + ".weight"]
    c3 = lambda n, t, s=1: F.conv3d(t, w(n), None, stride=s, padding=1)
    c1 = lambda n, t: F.conv3d(t, w(n), None)
    dp = (lambda i, t: t) if drop is None else (lambda i, t: t * drop[i][:t.shape[0]])
    out = c3("conv3d_c1_1", x)
    res = out
    out = c3("conv3d_c1_2", _lrelu(out))
    out = c3("lrelu_conv_c1.1", _lrelu(dp(0, out)))
    out = out + res
    ctx1 = _lrelu(out)
    out = _lrelu(_inorm(out))
    ctx = []
    for lvl in (2, 3, 4, 5):
        out = c3("conv3d_c%d" % lvl, out, 2)
        res = out
        name = "norm_lrelu_conv_c%d.2" % lvl
        out = c3(name, _lrelu(_inorm(out)))
        out = dp(lvl - 1, out)
        out = c3(name, _lrelu(_inorm(out)))
        out = out + res
        if lvl < 5:
            out = _lrelu(_inorm(out))
            ctx.append(out)
    ctx2, ctx3, ctx4 = ctx

    def up_block(name, t):  # norm_lrelu_upscale_conv_norm_lrelu (mask_branch.py:107-115)
        return _lrelu(_inorm(c3(name + ".3", _up(_lrelu(_inorm(t))))))

    out = up_block("norm_lrelu_upscale_conv_norm_lrelu_l0", out)
    out = _lrelu(_inorm(c1("conv3d_l0", out)))
    out = torch.cat([out, ctx4], 1)
    out = _lrelu(_inorm(c3("conv_norm_lrelu_l1.0", out)))
    out = c1("conv3d_l1", out)
    out = up_block("norm_lrelu_upscale_conv_norm_lrelu_l1", out)
    out = torch.cat([out, ctx3], 1)
    out = _lrelu(_inorm(c3("conv_norm_lrelu_l2.0", out)))
    ds2 = out
    out = c1("conv3d_l2", out)
    out = up_block("norm_lrelu_upscale_conv_norm_lrelu_l2", out)
    out = torch.cat([out, ctx2], 1)
    out = _lrelu(_inorm(c3("conv_norm_lrelu_l3.0", out)))
    ds3 = out
    out = c1("conv3d_l3", out)
    out = up_block("norm_lrelu_upscale_conv_norm_lrelu_l3", out)
    out = torch.cat([out, ctx1], 1)
    out = _lrelu(_inorm(c3("conv_norm_lrelu_l4.0", out)))
    pred = c1("conv3d_l4", out)
    s = _up(c1("ds2_1x1_conv3d", ds2)) + c1("ds3_1x1_conv3d", ds3)
    out = pred + _up(s)
    if stage == "finetune":
        out = _up(out) + F.conv3d(_up(out), w("out_upscale_conv.1"), None, padding=2)
    return out


# --------------------------------------------------------------------------------------
# losses  (reference model.py:808-1000)
# --------------------------------------------------------------------------------------


def rpn_class_loss(rpn_match, rpn_class_logits):
    """model.py:808-832.  rpn_match [A] in {-1,0,1}; logits [A,2]."""
    sel = torch.nonzero(rpn_match != 0)[:, 0]
    return F.cross_entropy(rpn_class_logits[sel], (rpn_match[sel] == 1).long())


def rpn_bbox_loss(target_bbox, rpn_match, rpn_bbox):
    """model.py:835-860.  target_bbox [T,6] zero padded; rpn_bbox [A,6]."""
    sel = torch.nonzero(rpn_match == 1)[:, 0]
    pred = rpn_bbox[sel]
    return F.smooth_l1_loss(pred, target_bbox[:pred.shape[0]])


def mrcnn_class_loss(target_class_ids, logits):
    """model.py:863-878 (called with class ids already collapsed to {0,1}, model.py:989)."""
    if target_class_ids.numel() == 0:
        return torch.zeros(1)
    return F.cross_entropy(logits, target_class_ids.long())


def mrcnn_bbox_loss(target_bbox, target_class_ids, pred_bbox):
    """model.py:881-906."""
    if target_class_ids.numel() == 0:
        return torch.zeros(1)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    cls = target_class_ids[pos].long()
    return F.smooth_l1_loss(pred_bbox[pos, cls], target_bbox[pos])


def mrcnn_mask_loss(target_masks, target_class_ids, pred_logits, class_weight=None):
    """model.py:909-935: argmax over the one-hot target, CrossEntropyLoss over logits."""
    if target_class_ids.numel() == 0:
        return torch.zeros(1)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    y_true = torch.argmax(target_masks[pos].long(), dim=1)
    return F.cross_entropy(pred_logits[pos], y_true, weight=class_weight)


def sobel_bank():
    """model.py:947-952."""
    kx = np.array([[[1, 2, 1], [0, 0, 0], [-1, -2, -1]],
                   [[2, 4, 2], [0, 0, 0], [-2, -4, -2]],
                   [[1, 2, 1], [0, 0, 0], [-1, -2, -1]]])
    return torch.from_numpy(np.array([kx, kx.transpose((1, 0, 2)), kx.transpose((0, 2, 1))])
                            .reshape((3, 1, 3, 3, 3))).float()


def mrcnn_mask_edge_loss(target_masks, target_class_ids, pred_masks):
    """model.py:938-981 incl. its quirks: magnitude uses response 0 twice and never 2; class loop is
    the literal range(7); target rows are taken as target_masks[:P, 1:]."""
    if target_class_ids.numel() == 0:
        return torch.zeros(1)
    kernel = sobel_bank().to(pred_masks.dtype)
    pos = torch.nonzero(target_class_ids > 0)[:, 0]
    P = pos.shape[0]
    y_true = target_masks[:P, 1:]
    y_pred = pred_masks[pos, 1:]
    loss = torch.zeros(1)
    for i in range(P):
        for j in range(7):
            gt = F.conv3d(y_true[i, j][None, None].to(pred_masks.dtype), kernel)
            gp = F.conv3d(y_pred[i, j][None, None], kernel)
            mt = torch.sqrt(gt[:, 0] ** 2 + gt[:, 1] ** 2 + gt[:, 0] ** 2)
            mp = torch.sqrt(gp[:, 0] ** 2 + gp[:, 1] ** 2 + gp[:, 0] ** 2)
            loss = loss + F.mse_loss(mp, mt)
    return loss / P


# --------------------------------------------------------------------------------------
# whole step  (reference model.py:1391-1514, 984-1000, 1632-1641)
# --------------------------------------------------------------------------------------


class Cfg(object):
    """The subset of HeartConfig (heart_main.py:26-174) the path reads, as plain attributes."""

    def __init__(self, image_dim=256, num_classes=8, stage="beginning", mask_pool=96, anchor_scales=(64, 128),
                 pool=12):
        self.IMAGE_SHAPE = (image_dim, image_dim, image_dim, 1)
        self.NUM_CLASSES = num_classes
        self.STAGE = stage
        self.BACKBONE_STRIDES = (8, 16)
        self.RPN_ANCHOR_SCALES = tuple(anchor_scales)
        self.RPN_BBOX_STD_DEV = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2)
        self.BBOX_STD_DEV = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2)
        self.PRE_NMS_LIMIT = 1000
        self.POST_NMS_ROIS_TRAINING = 500
        self.POST_NMS_ROIS_INFERENCE = 64
        self.RPN_NMS_THRESHOLD = 0.7
        self.TRAIN_ROIS_PER_IMAGE = 15
        self.ROI_POSITIVE_RATIO = 0.33
        self.DETECTION_TARGET_IOU_THRESHOLD = 0.5
        self.POOL_SIZE = (pool, pool, pool)
        self.MASK_POOL_SIZE = (mask_pool,) * 3
        self.MASK_SHAPE = tuple(2 * m for m in self.MASK_POOL_SIZE) if stage == "finetune" else self.MASK_POOL_SIZE
        self.DETECTION_MIN_CONFIDENCE = 0.7
        self.DETECTION_NMS_THRESHOLD = 0.3
        self.DETECTION_MAX_INSTANCES = 32
        self.LOSS_WEIGHTS = {"rpn_class_loss": 100., "rpn_bbox_loss": 50., "mrcnn_class_loss": 1.,
                             "mrcnn_bbox_loss": 20., "mrcnn_mask_loss": 1., "mrcnn_mask_edge_loss": 1.}

    def anchors(self):
        shapes = backbone_shapes(self.IMAGE_SHAPE, self.BACKBONE_STRIDES)
        return torch.from_numpy(generate_pyramid_anchors(self.RPN_ANCHOR_SCALES, shapes, self.BACKBONE_STRIDES, 1)).float()


def mold_image(vol):
    """model.py:1902-1904: (x - mean) / std with numpy's population std, on float32."""
    v = np.asarray(vol, dtype=np.float32)
    return (v - v.mean()) / v.std()


def train_forward(sd, cfg, image, rpn_match, rpn_bbox_t, gt_class_ids, gt_boxes, gt_masks,
                  perm_pos=None, perm_neg=None, drop=None, force_rois=None):
    """predict(mode='training') + compute_losses (model.py:1391-1514, 984-1000) for one volume.
    image [1,1,D,H,W] fp32 molded; rpn_match [A] int; rpn_bbox_t [T,6]; gt_boxes [G,6] pixels;
    gt_masks [8,D,H,W] fp32.  Returns dict with the six losses, the weighted total and intermediates."""
    H, W, D = cfg.IMAGE_SHAPE[:3]
    p2, p3 = fpn_forward(sd, image)
    lv = [rpn_forward(sd, p) for p in (p2, p3)]
    logits = torch.cat([l[0] for l in lv], 1)[0]
    probs = torch.cat([l[1] for l in lv], 1)[0]
    bbox = torch.cat([l[2] for l in lv], 1)[0]
    anchors = cfg.anchors()
    if force_rois is None:
        rois_n, keep, order = proposal_layer(probs, bbox, anchors, cfg.POST_NMS_ROIS_TRAINING, cfg.RPN_NMS_THRESHOLD,
                                             cfg.PRE_NMS_LIMIT, cfg.IMAGE_SHAPE, cfg.RPN_BBOX_STD_DEV)
    else:
        rois_n = force_rois
    scale = torch.tensor([D, H, W, D, H, W], dtype=torch.float32)
    p_rois, rois, tcls, tdel, tmask = detection_target_layer(
        rois_n, gt_class_ids, gt_boxes / scale, gt_masks, cfg.MASK_SHAPE, cfg.TRAIN_ROIS_PER_IMAGE,
        cfg.ROI_POSITIVE_RATIO, cfg.DETECTION_TARGET_IOU_THRESHOLD, cfg.BBOX_STD_DEV, perm_pos, perm_neg)
    out = {"rpn_class_logits": logits, "rpn_probs": probs, "rpn_bbox": bbox, "rpn_rois": rois_n, "p2": p2, "p3": p3,
           "rois": rois, "p_rois": p_rois, "target_class_ids": tcls, "target_deltas": tdel, "target_mask": tmask}
    zero = torch.zeros(1)
    l_rc = rpn_class_loss(rpn_match, logits)
    l_rb = rpn_bbox_loss(rpn_bbox_t, rpn_match, bbox)
    l_mc = l_mb = l_mm = l_me = zero
    if rois.shape[0] > 0:
        pooled = pyramid_roi_align(rois, [p2[0], p3[0]], cfg.POOL_SIZE)
        c_logits, c_probs, c_bbox = classifier_forward(sd, pooled)
        bin_ids = (tcls > 0).long()
        l_mc = mrcnn_class_loss(bin_ids, c_logits)
        l_mb = mrcnn_bbox_loss(tdel, bin_ids, c_bbox)
        out.update(mrcnn_class_logits=c_logits, mrcnn_bbox=c_bbox, pooled=pooled)
    if p_rois.shape[0] > 0:
        crops = pyramid_roi_align(p_rois, [image[0], image[0]], cfg.MASK_POOL_SIZE)
        m_logits = unet_forward(sd, crops, cfg.STAGE, drop)
        m_probs = F.softmax(m_logits, dim=1)
        l_mm = mrcnn_mask_loss(tmask, tcls, m_logits)
        if cfg.STAGE == "finetune":
            l_me = mrcnn_mask_edge_loss(tmask, tcls, m_probs)
        out.update(mask_crops=crops, mrcnn_mask_logits=m_logits, mrcnn_mask=m_probs)
    w = cfg.LOSS_WEIGHTS
    losses = [l_rc, l_rb, l_mc, l_mb, l_mm, l_me]
    total = (w["rpn_class_loss"] * l_rc + w["rpn_bbox_loss"] * l_rb + w["mrcnn_class_loss"] * l_mc +
             w["mrcnn_bbox_loss"] * l_mb + w["mrcnn_mask_loss"] * l_mm + w["mrcnn_mask_edge_loss"] * l_me)
    out["losses"] = losses
    out["loss"] = total
    return out


def train_step_grads(sd, cfg, trainable, *args, **kw):
    """Forward + backward + clip_grad_norm_(5.0) (model.py:1640-1641).  `trainable` lists the keys that
    receive gradients.  Returns (out dict, {key: grad}, pre-clip global norm)."""
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in trainable}
    sd2 = dict(sd)
    sd2.update(leaves)
    out = train_forward(sd2, cfg, *args, **kw)
    out["loss"].sum().backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    coef = min(1.0, 5.0 / (norm + 1e-6))
    grads = {k: g * coef for k, g in grads.items()}
    return out, grads, norm


# --------------------------------------------------------------------------------------
#  Inference pre / post-processing (model.py:1774-1864, utils.py:342-393, 443-460)
# --------------------------------------------------------------------------------------
def zoom_linear(vol, out_shape):
    """utils.resize(order=1, mode='constant') as skimage >= 0.19 evaluates it: scipy.ndimage.zoom(order=1,
    mode='grid-constant', cval=0, grid_mode=True) -- source coordinate (o + 0.5) * in / out - 0.5, linear blend of the two
    neighbours with everything outside the volume equal to 0, float64.  vol [H,W,D] -> out_shape (separable restatement)."""
    v = np.asarray(vol, dtype=np.float64)
    for ax, n_out in enumerate(out_shape):
        n_in = v.shape[ax]
        cc = (np.arange(n_out) + 0.5) * (n_in / float(n_out)) - 0.5
        i0 = np.floor(cc).astype(np.int64)
        t = cc - i0
        pad = np.concatenate([np.zeros_like(np.take(v, [0], axis=ax)), v, np.zeros_like(np.take(v, [0], axis=ax))], axis=ax)
        a = np.take(pad, np.clip(i0 + 1, 0, n_in + 1), axis=ax)
        b = np.take(pad, np.clip(i0 + 2, 0, n_in + 1), axis=ax)
        shp = [1] * v.ndim
        shp[ax] = n_out
        v = a * (1.0 - t).reshape(shp) + b * t.reshape(shp)
    return v


def mold_inputs(image_hwdc, min_dim, max_dim):
    """MaskRCNN.mold_inputs for IMAGE_RESIZE_MODE 'self' (model.py:1774-1810, utils.py:389-393): order-1 resize to
    [max, max, min], cast back to the image dtype (C truncation), (x - mean) / std, [H,W,D,C] -> [C,D,H,W]."""
    img = np.asarray(image_hwdc)
    r = zoom_linear(img[..., 0], (max_dim, max_dim, min_dim)).astype(img.dtype)[..., None]
    molded = mold_image(r).transpose((3, 2, 0, 1))
    window = (0, 0, 0, min_dim, max_dim, max_dim)
    return molded.astype(np.float32), window


def unmold_detections(detections, mask0_cdhw, image_shape_cdhw, window):
    """model.py:1812-1864 + utils.unmold_mask:443-460.  detections [n,8] (z1,y1,x1,z2,y2,x2,class,score), mask0 the class
    probabilities [ncls,d,h,w] of detection 0.  Returns (boxes [n,6] in (y1,x1,z1,y2,x2,z2), scores, full_mask [H,W,D])."""
    zero_ix = np.where(detections[:, 6] == 0)[0]
    N = zero_ix[0] if zero_ix.shape[0] > 0 else detections.shape[0]
    boxes = detections[:N, :6].astype(np.int32)
    scores = detections[:N, 7]
    sc = np.array([image_shape_cdhw[1] / (window[3] - window[0]), image_shape_cdhw[2] / (window[4] - window[1]),
                   image_shape_cdhw[3] / (window[5] - window[2])] * 2)
    sh = np.array(list(window[:3]) * 2)
    boxes = np.multiply(boxes - sh, sc).astype(np.int32)
    bad = np.where((boxes[:, 3] - boxes[:, 0]) * (boxes[:, 4] - boxes[:, 1]) * (boxes[:, 5] - boxes[:, 2]) <= 0)[0]
    if bad.shape[0] > 0:
        boxes, scores = np.delete(boxes, bad, 0), np.delete(scores, bad, 0)
        assert 0 not in bad, "restatement covers the case the goldens exercise: detection 0 survives"
    z1, y1, x1, z2, y2, x2 = [int(v) for v in boxes[0]]
    m = torch.from_numpy(np.ascontiguousarray(mask0_cdhw)).float().unsqueeze(0)
    m = F.interpolate(m, size=(z2 - z1, y2 - y1, x2 - x1), mode='trilinear', align_corners=False)[0].numpy()
    full = np.zeros((m.shape[0], image_shape_cdhw[1], image_shape_cdhw[2], image_shape_cdhw[3]), dtype=np.float32)
    full[:, z1:z2, y1:y2, x1:x2] = m
    full_mask = np.argmax(full, axis=0)
    boxes = boxes[:, [1, 2, 0, 4, 5, 3]]
    return boxes, scores, full_mask.transpose((1, 2, 0))
