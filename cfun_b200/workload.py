"""The synthetic workload of BASELINE.json config 2 (SURVEY.md 8d recipe), shared by BOTH bench arms and by the parity tests.

Pure numpy / torch-CPU and free of any import that loads libcfun_b200.so, so that the reference arm of bench.py
(`--impl reference`) can build bit-identical inputs and weights without mapping the product's native library into its
process.  Contents: int16 HU-like volume, label cube placement, the +5 % GT box of load_image_gt (reference
model.py:1058-1076), RPN targets (reference model.py:1090-1181), the initialize_weights recipe (model.py:1306-1319)."""
import math
import zlib

import numpy as np

WEIGHT_SEED = 2          # with this seed the untrained detector's proposals admit a 4-positive label placement
VOLUME_SEED = 1000
FWD_GFLOP = {"beginning": 1314.6, "finetune": 1767.6}      # H256, 4 positive / 12 RoIs
STEP_TFLOP = {"beginning": 3.94, "finetune": 5.30}


def workload_name(dim, stage):
    return ("MM-WHS-shape %d^3 synthetic int16 CT, 8-class heart, full train step (fwd + 6 losses + bwd + clip + SGD), stage %s, "
            "1 volume per step" % (dim, stage))


def shape_params(dim):
    """(mask pool, anchor scales, label cube side) of the cubic HeartConfig at `dim`"""
    mask_pool = 96 if dim >= 128 else 32
    scales = (64, 128) if dim >= 256 else ((32, 64) if dim >= 128 else (16, 32))
    return mask_pool, scales, 70 * dim // 256


def bench_weights(shapes, seed=WEIGHT_SEED):
    """MaskRCNN.initialize_weights recipe (reference model.py:1306-1319: xavier_uniform conv weights, zero conv bias,
    N(0, 0.01) linear weights, BN at identity) with a per-tensor generator keyed by (seed, name), so the GPU arm and the
    CPU reference arm build bit-identical weights from a seed alone."""
    import torch
    sd = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        g = torch.Generator().manual_seed((seed * 7919 + zlib.crc32(k.encode())) % (2 ** 31 - 1))
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
        elif k.endswith("running_var") or (k.endswith(".weight") and len(shp) == 1):
            sd[k] = torch.ones(shp)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            sd[k] = torch.zeros(shp)
        elif len(shp) == 5:
            rf = shp[2] * shp[3] * shp[4]
            bound = math.sqrt(6.0 / (shp[1] * rf + shp[0] * rf))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.01
    return sd


def build_rpn_targets(anchors, gt_boxes, config):
    """anchor/GT matching and delta targets (reference model.py:1090-1181), vectorised numpy."""
    n_t = config.RPN_TRAIN_ANCHORS_PER_IMAGE
    rpn_match = np.zeros([anchors.shape[0]], dtype=np.int32)
    rpn_bbox = np.zeros((n_t, 6))
    va = np.prod(anchors[:, 3:] - anchors[:, :3], axis=1)
    overlaps = np.zeros((anchors.shape[0], gt_boxes.shape[0]))
    for j, g in enumerate(gt_boxes):
        lo = np.maximum(anchors[:, :3], g[:3])
        hi = np.minimum(anchors[:, 3:], g[3:])
        inter = np.prod(np.maximum(hi - lo, 0)[:, ::-1], axis=1)
        overlaps[:, j] = inter / (np.prod(g[3:] - g[:3]) + va - inter + 1e-6)
    amax = np.argmax(overlaps, axis=1)
    vmax = overlaps[np.arange(overlaps.shape[0]), amax]
    rpn_match[vmax < 0.3] = -1
    rpn_match[np.argmax(overlaps, axis=0)] = 1
    rpn_match[vmax >= 0.7] = 1
    ids = np.where(rpn_match == 1)[0]
    extra = len(ids) - n_t // 2
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == -1)[0]
    extra = len(ids) - (n_t - np.sum(rpn_match == 1))
    if extra > 0:
        rpn_match[np.random.choice(ids, extra, replace=False)] = 0
    ids = np.where(rpn_match == 1)[0]
    a = anchors[ids]
    g = gt_boxes[amax[ids]]
    asz, gsz = a[:, 3:] - a[:, :3], g[:, 3:] - g[:, :3]
    actr, gctr = a[:, :3] + 0.5 * asz, g[:, :3] + 0.5 * gsz
    tgt = np.concatenate([(gctr - actr) / asz, np.log(gsz / asz)], axis=1) / np.asarray(config.RPN_BBOX_STD_DEV)
    rpn_bbox[:len(ids)] = tgt[:n_t]
    return rpn_match, rpn_bbox


def synth_volume(dim, seed, cube=70):
    rng = np.random.default_rng(seed)
    vol = np.clip(np.round(rng.standard_normal((dim, dim, dim), dtype=np.float32) * 300.0), -1024, 3071).astype(np.int16)
    lab = np.zeros((dim, dim, dim), dtype=np.uint8)        # [H,W,D]
    a = (dim - cube) // 2
    lab[a:a + cube, a:a + cube, a:a + cube] = rng.integers(1, 8, size=(cube, cube, cube), dtype=np.uint8)
    return vol, lab


def gt_box_from_label(lab_hwd, num_classes):
    """bbox of the labelled region in (z,y,x) order with the reference's 5 % margin, tiled NUM_CLASSES-1 times."""
    lab = lab_hwd.transpose((2, 0, 1))
    nz = np.nonzero(lab)
    lo = np.array([v.min() for v in nz], dtype=np.float64)
    hi = np.array([v.max() + 1 for v in nz], dtype=np.float64)
    ext = hi - lo
    lo = np.floor(np.maximum(0, lo - 0.05 * ext))
    hi = np.ceil(np.minimum(lab.shape, hi + 0.05 * ext))
    box = np.concatenate([lo, hi]).astype(np.int32)
    return np.tile(box[None], (num_classes - 1, 1))


# ---------------------------------------------------------------------------------------------------------
# label placement that satisfies the benchmark precondition (4 positive / 12 sampled RoIs)
# ---------------------------------------------------------------------------------------------------------
def _iou_many(boxes, cands):
    """IoU of every candidate [m,6] against every box [n,6] (pixels) -> [m,n]"""
    lo = np.maximum(cands[:, None, :3], boxes[None, :, :3])
    hi = np.minimum(cands[:, None, 3:], boxes[None, :, 3:])
    inter = np.prod(np.clip(hi - lo, 0, None), axis=2)
    vc = np.prod(cands[:, 3:] - cands[:, :3], axis=1)[:, None]
    vb = np.prod(boxes[:, 3:] - boxes[:, :3], axis=1)[None, :]
    return inter / (vc + vb - inter + 1e-9)


def _final_gt_box(center, side, dim):
    """label cube (start, L) and the GT box load_image_gt derives from it (+5 % margin, floor / ceil, clipped).
    dim: scalar (cubic volume) or (D, H, W)"""
    dim = np.broadcast_to(np.asarray(dim), (3,))
    L = int(round(side / 1.1))
    start = np.clip(np.round(center - L / 2.0).astype(int), 0, dim - L)
    lo = np.floor(np.maximum(0, start - 0.05 * L))
    hi = np.ceil(np.minimum(dim, start + L + 0.05 * L))
    return start, L, np.concatenate([lo, hi]).astype(np.float64)


def place_label_cube(rois_norm, dim, want=4, sides=(72, 80, 88, 96, 104), margin=0.03):
    """With random-init weights and a noise volume the RPN's proposals are unrelated to any fixed label, so a centred
    cube usually yields zero positive RoIs and the U-Net (92 % of the step's FLOPs) never runs.  The benchmark therefore
    places the synthetic label cube where the untrained detector's proposals cluster: the cube whose GT box (as
    load_image_gt derives it, reference model.py:1058-1075) has the most proposals with IoU >= 0.5 (at least `want`, none
    within `margin` of the threshold).  Returns (start_zyx, side, n_positive_candidates) or None."""
    d3 = np.broadcast_to(np.asarray(dim, dtype=np.float64), (3,))
    boxes = np.asarray(rois_norm, dtype=np.float64) * np.concatenate([d3, d3])
    ctr = 0.5 * (boxes[:, :3] + boxes[:, 3:])
    cents = [ctr]
    d2 = ((ctr[:, None, :] - ctr[None, :, :]) ** 2).sum(-1)
    nn = np.argsort(d2, axis=1)
    for k in (2, 4, 8):
        cents.append(ctr[nn[:, :k]].mean(axis=1))
    cents = np.unique(np.round(np.concatenate(cents, 0)), axis=0)
    best = None
    for s in sides:
        finals = [_final_gt_box(c, s, dim) for c in cents]
        cand = np.stack([f[2] for f in finals])
        iou = _iou_many(boxes, cand)
        npos = (iou >= 0.5 + margin).sum(1)
        amb = ((iou > 0.5 - margin) & (iou < 0.5 + margin)).sum(1)
        ok = (npos >= want) & (amb == 0)
        if not ok.any():
            continue
        score = np.where(ok, npos + iou.max(1) * 0.5, -1)
        i = int(np.argmax(score))
        if best is None or score[i] > best[0]:
            best = (score[i], finals[i][0], finals[i][1], int(npos[i]))
    if best is None:
        return None
    return best[1], best[2], best[3]


def label_from_cube(dim, start_zyx, side, seed, num_classes=8):
    """uint8 label volume [H,W,D] with a cube of uniformly random classes 1..num_classes-1 at (z,y,x) = start.
    dim: scalar (cubic volume) or (D, H, W)"""
    rng = np.random.default_rng(seed)
    D, H, W = [int(v) for v in np.broadcast_to(np.asarray(dim), (3,))]
    lab = np.zeros((H, W, D), dtype=np.uint8)      # [H,W,D]
    z, y, x = [int(v) for v in start_zyx]
    lab[y:y + side, x:x + side, z:z + side] = rng.integers(1, num_classes, size=(side, side, side), dtype=np.uint8)
    return lab
