"""TEST INFRASTRUCTURE: golden vectors for BASELINE config 1 (single 64^3 volume, forward-only, `heart_main.py test` path)
from the UNMODIFIED reference: MaskRCNN.detect() = mold_inputs (resize + normalise) -> predict('inference') ->
unmold_detections (box rescale, trilinear unmold_mask, argmax).  Run in the build container (needs /root/reference):

    python oracle/gen_golden_inference.py        ->  tests/golden/inference64.npz

With random-init weights no RoI passes class > 0 & score >= 0.7 and the reference dies; the classifier's
class bias is therefore set to (-1, +1) -- "votes foreground" -- so that detections survive; everything else is the
deterministic weight recipe of oracle/detweights.py (seed 200)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim                      # noqa: E402
from detweights import det_state    # noqa: E402
from gen_golden import make_config, save      # noqa: E402


def main():
    torch.set_num_threads(8)
    R = refshim.load()
    M = R["model"]
    cfg = make_config(R, 64, "beginning", mask_pool=32, anchor_scales=(16, 32))
    torch.manual_seed(5)
    with refshim.quiet():
        net = M.MaskRCNN(cfg, "/tmp/_cfun_golden", test_flag=True)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = det_state(shapes, seed=200)
    sd["classifier.linear_class.bias"] = torch.tensor([-1.0, 1.0])
    net.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(2024)
    H, W, D = 80, 72, 48                                   # raw scan, resized to 64^3 by mold_inputs ('self' mode)
    vol = np.clip(np.round(rng.normal(0, 300, size=(H, W, D))), -1024, 3071).astype(np.int16)
    image = vol[..., None]                                 # [H,W,D,1] as heart_main.py:303 builds it
    with refshim.quiet():
        molded, metas, windows = net.mold_inputs([image])
        with torch.no_grad():
            det, mmask = net.predict([torch.from_numpy(molded).float(), metas], mode="inference")
        res = net.detect([image])[0]
    det = det.detach().numpy()[0]
    mm = mmask.detach().numpy()[0]                         # [n, 8, 32, 32, 32]
    print("detections", det.shape, "first", det[0], "mask classes present", np.unique(res["mask"]))
    save("inference64", vol=vol, molded=molded[0].astype(np.float32), image_meta=metas[0].astype(np.float64),
         window=windows[0].astype(np.int64), detections=det.astype(np.float32), mask0=mm[0].astype(np.float32),
         mask_sample=mm.reshape(-1)[::1009].astype(np.float32), rois=res["rois"].astype(np.int32),
         class_ids=np.asarray(res["class_ids"]).astype(np.int32), scores=np.asarray(res["scores"]).astype(np.float32),
         full_mask=res["mask"].astype(np.uint8))
    # float64 yardstick of the mask probabilities (the oracle restatement in double on the reference's own crops)
    sys.path.insert(0, HERE)
    import cfun_oracle as O
    img = torch.from_numpy(molded[0].astype(np.float32))[None]
    boxes = torch.from_numpy(det[:, :6] / 64.0).float()
    crops = O.pyramid_roi_align(boxes, [img[0], img[0]], (32, 32, 32))
    sd64 = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in sd.items()}
    with torch.no_grad():
        y64 = torch.softmax(O.unet_forward(sd64, crops.double(), "beginning", None), 1)
    save("inference64_fp64", mask0_fp64=y64[0].numpy(), mask_sample_fp64=y64.reshape(-1)[::1009].numpy())


if __name__ == "__main__":
    main()
