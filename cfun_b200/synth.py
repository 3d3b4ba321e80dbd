"""Synthetic CT volumes and per-step training inputs (SURVEY.md 8d recipe): int16 HU-like volume, a centred label cube
with uniformly random classes 1..7, the +5 % GT box of load_image_gt (reference model.py:1058-1076) and RPN targets.
Host-side (numpy); the device work starts at MaskRCNN.train_step_from_host."""
import numpy as np
import torch

from . import model as M


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


class StepInputs(object):
    """Pinned host buffers of one training step (what the H2D copy moves) + their byte count."""

    def __init__(self, cfg, anchors_np, dim, seed, cube=70, pin=True, vol=None, lab=None):
        if vol is None:
            vol, lab = synth_volume(dim, seed, cube)
        boxes = gt_box_from_label(lab, cfg.NUM_CLASSES)
        state = np.random.get_state()
        np.random.seed(seed % (2 ** 31))
        rpn_match, rpn_bbox = M.build_rpn_targets(anchors_np, boxes[:1].astype(np.float32), cfg)
        np.random.set_state(state)
        mk = (lambda t: t.pin_memory()) if (pin and torch.cuda.is_available()) else (lambda t: t)
        self.vol = mk(torch.from_numpy(vol))                                     # int16 [H,W,D]
        self.label = mk(torch.from_numpy(lab))                                   # uint8 [H,W,D]
        self.rpn_match = mk(torch.from_numpy(rpn_match.astype(np.int32)))        # [A]
        self.rpn_bbox = mk(torch.from_numpy(rpn_bbox.astype(np.float32)))        # [T,6]
        self.gt_boxes = mk(torch.from_numpy(boxes.astype(np.float32)))           # [7,6]
        self.gt_class_ids = mk(torch.arange(1, cfg.NUM_CLASSES, dtype=torch.int32))

    def tensors(self):
        return (self.vol, self.label, self.rpn_match, self.rpn_bbox, self.gt_boxes, self.gt_class_ids)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())


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
    """label cube (start, L) and the GT box load_image_gt derives from it (+5 % margin, floor / ceil, clipped)"""
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
    boxes = np.asarray(rois_norm, dtype=np.float64) * dim
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


def label_from_cube(dim, start_zyx, side, seed):
    """uint8 label volume [H,W,D] with a cube of uniformly random classes 1..7 at (z,y,x) = start"""
    rng = np.random.default_rng(seed)
    lab = np.zeros((dim, dim, dim), dtype=np.uint8)      # [H,W,D]
    z, y, x = [int(v) for v in start_zyx]
    lab[y:y + side, x:x + side, z:z + side] = rng.integers(1, 8, size=(side, side, side), dtype=np.uint8)
    return lab
