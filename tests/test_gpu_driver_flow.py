"""GPU: the public driver flow of heart_main.py (train(): Dataset -> MaskRCNN.train_model; test(): MaskRCNN.detect) on synthetic
NIfTI volumes, through the same calls the reference driver makes (heart_main.py:264-283, 286-330)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_model_and_detect_on_synthetic_nifti(tmp_path, monkeypatch):
    from cfun_b200 import config as Cf, model as M, utils as U, nifti
    nifti.install_as_nibabel()
    import nibabel as nib
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(5)
    paths = []
    for i in range(2):
        vol = np.clip(np.round(rng.normal(0, 300, size=(64, 64, 64))), -1024, 3071).astype(np.int16)
        lab = np.zeros((64, 64, 64), dtype=np.int16)
        lab[16:44, 20:48, 18:46] = rng.integers(1, 8, size=(28, 28, 28))
        nib.save(nib.Nifti1Image(vol, np.eye(4)), str(tmp_path / ("img%d.nii.gz" % i)))
        nib.save(nib.Nifti1Image(lab, np.eye(4)), str(tmp_path / ("lab%d.nii.gz" % i)))
        paths.append((str(tmp_path / ("img%d.nii.gz" % i)), str(tmp_path / ("lab%d.nii.gz" % i))))

    class HeartLike(U.Dataset):          # what heart_main.HeartDataset does, on top of the same base class
        def load(self):
            for k in range(1, 8):
                self.add_class("heart", k, "c%d" % k)
            for img, lab in paths:
                self.add_image("heart", image_id=img, path=img, mask=lab)

        def load_mask(self, image_id):
            return nib.load(self.image_info[image_id]["mask"]).get_data().copy()

        def process_mask(self, mask):
            masks = np.stack([(mask == c) for c in range(self.num_classes)]).astype(np.int32)
            return masks, np.arange(1, self.num_classes, dtype=np.int32)

    ds = HeartLike()
    ds.load()
    ds.prepare()
    cfg = Cf.heart_config(64, "beginning", mask_pool=32, anchor_scales=(16, 32), STEPS_PER_EPOCH=2, LOADER_WORKERS=0,
                          DETECTION_MIN_CONFIDENCE=0.0)
    net = M.MaskRCNN(config=cfg, model_dir=str(tmp_path / "logs"), test_flag=False)
    net = net.cuda()
    before = {k: v.detach().clone() for k, v in net.state_dict().items() if k in ("rpn.conv_shared.weight", "fpn.P2_conv2.weight")}
    net.train_model(ds, ds, learning_rate=cfg.LEARNING_RATE, epochs=1)
    torch.cuda.synchronize()
    after = net.state_dict()
    assert all(torch.isfinite(after[k]).all() for k in before)
    assert any(not torch.equal(after[k], before[k]) for k in before), "the optimizer step changed nothing"
    # inference through detect(): make the binary head vote foreground so that detections survive
    with torch.no_grad():
        net.classifier.linear_class.bias.copy_(torch.tensor([-1.0, 1.0]))
    image = np.expand_dims(nib.load(paths[0][0]).get_data().copy(), -1)
    r = net.detect([image])[0]
    assert set(r) == {"rois", "class_ids", "scores", "mask"}
    assert r["mask"].shape == (64, 64, 64) and r["rois"].shape[1] == 6 and r["mask"].min() >= 0 and r["mask"].max() <= 7
