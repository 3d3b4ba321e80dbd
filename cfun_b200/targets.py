"""Device-side training-target generation: what reference load_image_gt / build_rpn_targets
(model.py:1007-1087, 1090-1181) compute with numpy on the host for every step -- the GT box of the label volume with its
5 % margin, the anchor/GT IoU match (+1 / -1 / 0) with random sub-sampling, and the delta targets of the positive anchors --
as static-shape device ops: no host round trip, no np.random, no data-dependent shapes (the random sub-sampling draws are
per-anchor keys that can be injected, so tests reproduce the reference's np.random.choice picks exactly).

IoU and the box refinement are the library's kernels (cfun_iou3d_eps: utils.compute_iou's unfused fp32 arithmetic incl. its
1e-6 epsilon; cfun_box_refinement); the rest is index algebra over the [A] anchor axis (A = 36 864 at 256^3)."""
import torch

from . import ops


def gt_box_from_label(label_dhw, num_classes):
    """load_image_gt's box (model.py:1058-1076): bounding box of the labelled voxels in (z,y,x), grown by 5 % of its extent,
    floor / ceil, clipped, tiled NUM_CLASSES-1 times.  label [D,H,W] integer class ids on device -> float32 [ncls-1, 6]."""
    fg = label_dhw > 0
    D, H, W = fg.shape
    lo, hi = [], []
    for ax, n in ((0, D), (1, H), (2, W)):
        other = tuple(a for a in (0, 1, 2) if a != ax)
        line = fg.amax(dim=other).to(torch.int32)                   # [n] 1 where the slab holds a labelled voxel
        first = torch.argmax(line)                                  # first 1
        last = n - 1 - torch.argmax(torch.flip(line, (0,)))         # last 1
        lo.append(first)
        hi.append(last + 1)
    lo = torch.stack(lo).double()
    hi = torch.stack(hi).double()
    ext = hi - lo
    size = torch.tensor([D, H, W], dtype=torch.float64, device=fg.device)
    lo = torch.floor(torch.clamp(lo - 0.05 * ext, min=0))
    hi = torch.ceil(torch.minimum(size, hi + 0.05 * ext))
    box = torch.cat([lo, hi]).to(torch.int32).float()
    return box.unsqueeze(0).repeat(num_classes - 1, 1)


def _reset_extra(match, value, limit, keys):
    """np.random.choice(ids, extra, replace=False) -> 0 for the anchors with match == value beyond `limit` (a device scalar
    or int): among those anchors the `extra` = count - limit with the smallest keys are reset to neutral."""
    member = match == value
    k = torch.where(member, keys, torch.full_like(keys, float("inf")))
    order = torch.argsort(k, stable=True)                           # members first, ascending key
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=order.device)
    count = member.sum()
    extra = torch.clamp(count - limit, min=0)
    return torch.where(member & (rank < extra), torch.zeros_like(match), match)


def build_rpn_targets(anchors, gt_boxes, config, keys_pos=None, keys_neg=None, generator=None):
    """reference model.py:1090-1181 on device.  anchors [A,6] fp32 (pixels), gt_boxes [G,6] fp32 -> (rpn_match int32 [A],
    rpn_bbox fp32 [RPN_TRAIN_ANCHORS_PER_IMAGE, 6]).  keys_pos / keys_neg: per-anchor sub-sampling keys [A] (uniform
    random if omitted, drawn on the device from `generator`)."""
    A = anchors.shape[0]
    dev = anchors.device
    n_t = int(config.RPN_TRAIN_ANCHORS_PER_IMAGE)
    overlaps = torch.stack([ops.iou_with_eps(gt_boxes[j:j + 1], anchors)[0] for j in range(gt_boxes.shape[0])], dim=1)   # [A,G]
    iou_max, iou_arg = overlaps.max(dim=1)
    match = torch.zeros(A, dtype=torch.int32, device=dev)
    match = torch.where(iou_max < 0.3, torch.full_like(match, -1), match)
    match[torch.argmax(overlaps, dim=0)] = 1                        # every GT box keeps its best anchor
    match = torch.where(iou_max >= 0.7, torch.ones_like(match), match)
    if keys_pos is None:
        keys_pos = torch.rand(A, device=dev, generator=generator)
    if keys_neg is None:
        keys_neg = torch.rand(A, device=dev, generator=generator)
    match = _reset_extra(match, 1, n_t // 2, keys_pos.float())
    match = _reset_extra(match, -1, n_t - (match == 1).sum(), keys_neg.float())
    # delta targets of the positive anchors, in ascending anchor order, rows beyond their count stay zero
    pos = match == 1
    row = torch.cumsum(pos.to(torch.int64), 0) - 1                  # destination row of every positive anchor
    deltas = ops.box_refinement(anchors, gt_boxes[iou_arg], config.RPN_BBOX_STD_DEV)          # [A,6]; used where pos
    rpn_bbox = torch.zeros((n_t + 1, 6), device=dev)
    dst = torch.where(pos & (row < n_t), row, torch.full_like(row, n_t))                      # non-positives -> scratch row
    rpn_bbox.index_copy_(0, dst, torch.where(pos.unsqueeze(1), deltas, torch.zeros_like(deltas)))
    return match, rpn_bbox[:n_t]
