"""TEST INFRASTRUCTURE: rebuilds the tensors of the golden 64^3 step from the stored raw volume/label."""
import numpy as np
import torch
import cfun_oracle as O


def golden_step_inputs(g):
    vol, lab = g["vol"], g["label"]                                # [H,W,D]
    image = O.mold_image(vol.astype(np.float32)[..., None]).transpose((3, 2, 0, 1))[None]   # model.py:1055,1085
    labt = lab.transpose((2, 0, 1))
    nz = np.argwhere(labt > 0)
    lo, hi = nz.min(0), nz.max(0) + 1
    gt_box = np.array([[lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]]], dtype=np.float32)
    return {
        "image": torch.from_numpy(np.ascontiguousarray(image)).float(),
        "label_dhw": torch.from_numpy(np.ascontiguousarray(labt)),
        "gt_boxes": torch.from_numpy(np.tile(gt_box, (7, 1))),
        "gt_masks": torch.from_numpy(np.stack([(labt == c) for c in range(8)]).astype(np.float32)),
        "rpn_match": torch.from_numpy(g["rpn_match"].astype(np.int32)),
        "rpn_bbox": torch.from_numpy(g["rpn_bbox"]).float(),
        "drop": [torch.from_numpy(g["drop%d" % i]) for i in range(5)],
    }
