"""Per-step training inputs as pinned host buffers (what the H2D copy of MaskRCNN.train_step_from_host moves).
The recipe itself (volume, label placement, GT box, RPN targets) lives in cfun_b200.workload."""
import numpy as np
import torch

from .workload import (synth_volume, gt_box_from_label, place_label_cube, label_from_cube, build_rpn_targets)   # noqa: F401


class StepInputs(object):
    """Pinned host buffers of one training step (what the H2D copy moves) + their byte count."""

    def __init__(self, cfg, anchors_np, dim, seed, cube=70, pin=True, vol=None, lab=None):
        if vol is None:
            vol, lab = synth_volume(dim, seed, cube)
        boxes = gt_box_from_label(lab, cfg.NUM_CLASSES)
        state = np.random.get_state()
        np.random.seed(seed % (2 ** 31))
        rpn_match, rpn_bbox = build_rpn_targets(anchors_np, boxes[:1].astype(np.float32), cfg)
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
