"""CPU-only tests: C-ABI library loads and exports every declared symbol, checkpoint ABI, host-side logic, and the
world_size-2 (gloo) data-parallel plumbing.  No CUDA compute is launched here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import cfun_oracle as O
from conftest import load_golden, ROOT
from shapes import maskrcnn_shapes


def header_functions():
    txt = open(os.path.join(ROOT, "include", "cfun_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cfun_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from cfun_b200 import _lib
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(_lib.lib, n), "libcfun_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert _lib.lib.cfun_version() >= 100


def test_sass_is_sm100a_only():
    from cfun_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_checkpoint_abi_matches_reference_table():
    from cfun_b200 import model as M, config as Cf
    cfg = Cf.heart_config(64, "beginning", mask_pool=32, anchor_scales=(16, 32))
    net = M.MaskRCNN(cfg, "/tmp/_cfun_test")
    sd = net.state_dict()
    want = maskrcnn_shapes()
    assert len(sd) == 220 and set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    trainable = sum(p.numel() for p in net.parameters() if p.requires_grad)
    assert trainable == 41349586
    assert all(not p.requires_grad for n, p in net.named_parameters() if ".bn" in n or ".C1.1." in n or "downsample.1" in n)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cfun_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cfun_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
    for shim in ("model.py", "utils.py", "backbone.py", "mask_branch.py", "config.py"):
        assert "oracle" not in open(os.path.join(ROOT, shim)).read()


def test_root_shims_expose_reference_names():
    import model, utils, backbone, mask_branch, config
    for name in ("MaskRCNN", "FPN", "RPN", "Classifier", "Mask", "proposal_layer", "RoI_Align", "pyramid_roi_align",
                 "detection_target_layer", "refine_detections", "detection_layer", "apply_box_deltas", "clip_boxes",
                 "bbox_overlaps", "compute_losses", "compute_mrcnn_mask_edge_loss", "build_rpn_targets", "mold_image",
                 "compose_image_meta", "parse_image_meta", "compute_backbone_shapes", "Dataset", "log"):
        assert hasattr(model, name), name
    for name in ("non_max_suppression", "compute_iou", "box_refinement", "generate_anchors", "generate_pyramid_anchors",
                 "denorm_boxes_graph", "Dataset", "compute_per_class_mask_iou", "extract_bboxes", "resize_image", "unmold_mask"):
        assert hasattr(utils, name), name
    assert hasattr(backbone, "P3D19") and hasattr(backbone, "Bottleneck") and hasattr(backbone, "P3D")
    assert hasattr(mask_branch, "Modified3DUNet")
    assert hasattr(config, "Config")


def test_anchor_order_matches_reference_golden():
    from cfun_b200 import utils as U, model as M
    g = load_golden("anchors")

    class C: BACKBONE_STRIDES = list(g["strides"])
    shapes = M.compute_backbone_shapes(C, tuple(g["image_shape"]))
    assert np.array_equal(shapes, g["shapes"])
    a = U.generate_pyramid_anchors(tuple(g["scales"]), [1], shapes, list(g["strides"]), 1)
    assert np.array_equal(a, g["anchors"])


def test_rpn_targets_match_reference_golden():
    from cfun_b200 import model as M, config as Cf
    g = load_golden("step64_beginning")
    cfg = Cf.heart_config(64, "beginning", mask_pool=32, anchor_scales=(16, 32))
    anchors = O.generate_pyramid_anchors((16, 32), O.backbone_shapes((64, 64, 64, 1), (8, 16)), (8, 16)).astype(np.float32)
    lab = g["label"].transpose((2, 0, 1))
    nz = np.argwhere(lab > 0)
    gt = np.concatenate([nz.min(0), nz.max(0) + 1])[None].astype(np.int32)
    np.random.seed(9)
    match, bbox = M.build_rpn_targets(anchors, gt, cfg)
    assert np.array_equal(match, g["rpn_match"])
    assert np.allclose(bbox, g["rpn_bbox"], rtol=1e-6, atol=1e-7)


def test_config_is_name_compatible():
    from cfun_b200 import config as Cf
    c = Cf.HeartConfig("finetune")
    assert tuple(c.IMAGE_SHAPE) == (320, 320, 192, 1) and c.MASK_SHAPE == (192, 192, 192) and c.BATCH_SIZE == 1
    assert Cf.HeartConfig("beginning").MASK_SHAPE == (96, 96, 96)
    for name in ("PRE_NMS_LIMIT", "UNET_MASK_BRANCH_CHANNEL", "RPN_BBOX_STD_DEV", "DETECTION_TARGET_IOU_THRESHOLD",
                 "LOSS_WEIGHTS", "TOP_DOWN_PYRAMID_SIZE", "RPN_CONV_CHANNELS", "POST_NMS_ROIS_TRAINING"):
        assert hasattr(c, name)


def test_volume_sharding_is_a_partition():
    from cfun_b200.dp import shard_volumes
    for world in (1, 2, 4, 8):
        got = sorted(sum((shard_volumes(8, r, world) for r in range(world)), []))
        assert got == list(range(8))


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from cfun_b200.dp import flatten_grads_cpu
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["MASTER_PORT"], rank=rank, world_size=world)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Conv3d(2, 4, 3, padding=1), torch.nn.Conv3d(4, 2, 1))
x = torch.randn(1, 2, 6, 6, 6, generator=torch.Generator().manual_seed(100 + rank))       # one volume per rank
net(x).square().sum().backward()
local = [p.grad.clone() for p in net.parameters()]
flat = flatten_grads_cpu(list(net.parameters()))
dist.all_reduce(flat, op=dist.ReduceOp.SUM)                                                  # the one collective
# every rank recomputes the other ranks' gradients serially and checks allreduce == sum of per-volume gradients
tot = [torch.zeros_like(g) for g in local]
for r in range(world):
    net.zero_grad(set_to_none=True)
    xr = torch.randn(1, 2, 6, 6, 6, generator=torch.Generator().manual_seed(100 + r))
    net(xr).square().sum().backward()
    for t, p in zip(tot, net.parameters()):
        t += p.grad
off = 0
for t in tot:
    k = t.numel()
    assert torch.allclose(flat[off:off + k].view_as(t), t, rtol=1e-5, atol=1e-6), "allreduce != sum of per-volume grads"
    off += k
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = str(29500 + os.getpid() % 500)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


def test_nifti_round_trip_and_reference_surface(tmp_path):
    """cfun_b200.nifti: the slice of nibabel the reference driver uses (heart_main.py:211,300-303,349-352)"""
    import gzip
    import struct
    from cfun_b200 import nifti
    rng = np.random.default_rng(0)
    aff = np.array([[0.4, 0, 0, -100.0], [0, 0.5, 0, -90.0], [0, 0, 0.8, 30.0], [0, 0, 0, 1.0]])
    for dt, ext in ((np.int16, ".nii.gz"), (np.int32, ".nii"), (np.float32, ".nii.gz"), (np.uint8, ".nii")):
        vol = (rng.normal(0, 300, size=(7, 9, 5))).astype(dt)
        p = str(tmp_path / ("v" + ext))
        nifti.save(nifti.Nifti1Image(vol, aff), p)
        img = nifti.load(p)
        got = img.get_data().copy()
        assert got.dtype == vol.dtype and got.shape == (7, 9, 5) and np.array_equal(got, vol)
        assert np.allclose(img.affine, aff) and np.allclose(img.header.get_zooms(), (0.4, 0.5, 0.8))
        assert img.get_fdata().dtype == np.float64
    # a big-endian file with scl_slope / scl_inter and a qform-only affine, written by hand
    vol = np.arange(24, dtype=">i2").reshape((2, 3, 4), order="F")
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">2h", hdr, 70, 4, 16)
    struct.pack_into(">8f", hdr, 76, 1.0, 2.0, 3.0, 4.0, 1, 1, 1, 1)
    struct.pack_into(">3f", hdr, 108, 352.0, 0.5, 10.0)
    struct.pack_into(">2h", hdr, 252, 1, 0)
    struct.pack_into(">6f", hdr, 256, 0.0, 0.0, 0.0, 5.0, 6.0, 7.0)
    hdr[344:348] = b"n+1\0"
    p = str(tmp_path / "be.nii.gz")
    with gzip.open(p, "wb") as f:
        f.write(bytes(hdr) + b"\0" * 4 + vol.tobytes(order="F"))
    img = nifti.load(p)
    assert np.allclose(img.get_data(), np.arange(24).reshape((2, 3, 4), order="F") * 0.5 + 10.0)
    assert np.allclose(img.affine, np.array([[2.0, 0, 0, 5], [0, 3.0, 0, 6], [0, 0, 4.0, 7], [0, 0, 0, 1]]))
    with pytest.raises(ValueError):
        (tmp_path / "junk.nii").write_bytes(b"x" * 400)
        nifti.load(str(tmp_path / "junk.nii"))


@pytest.mark.skipif(not os.path.exists("/root/reference/heart_main.py"), reason="reference tree not present")
def test_reference_driver_imports_against_the_shims():
    """heart_main.py, unmodified, resolves `config`, `model`, `utils` (and nibabel) to this repository: its Config /
    Dataset subclasses build on our base classes and its derived shapes match the reference's"""
    import importlib.util
    from cfun_b200 import nifti
    nifti.install_as_nibabel()
    saved = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        spec = importlib.util.spec_from_file_location("heart_main_ref", "/root/reference/heart_main.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.dont_write_bytecode = saved
    assert m.HeartConfig.__mro__[1].__module__ == "cfun_b200.config"
    assert m.HeartDataset.__mro__[1].__module__ == "cfun_b200.utils"
    cfg = m.HeartConfig("beginning")
    assert list(cfg.IMAGE_SHAPE) == [320, 320, 192, 1] and tuple(cfg.MASK_SHAPE) == (96, 96, 96) and cfg.NUM_CLASSES == 8
    assert tuple(m.HeartConfig("finetune").MASK_SHAPE) == (192, 192, 192)
    ds = m.HeartDataset()
    assert hasattr(ds, "add_class") and hasattr(ds, "prepare") and callable(m.train) and callable(m.test)


@pytest.mark.skipif(not os.path.exists("/root/reference/heart_main.py"), reason="reference tree not present")
def test_reference_dataset_feeds_our_data_layer(tmp_path):
    """The reference's own HeartDataset (heart_main.py:181-261, unmodified) reading synthetic NIfTI volumes through the
    nibabel stand-in, fed to this repository's data layer (model.Dataset -> load_image_gt): the host side of
    `heart_main.py train` up to the H2D copy."""
    import importlib.util
    import json
    import types
    from cfun_b200 import nifti, model as M, utils as U
    nifti.install_as_nibabel()
    saved = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        spec = importlib.util.spec_from_file_location("heart_main_ref2", "/root/reference/heart_main.py")
        hm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(hm)
    finally:
        sys.dont_write_bytecode = saved
    rng = np.random.default_rng(3)
    entries = []
    for i in range(14):                       # load_heart() keeps entries [13:] for 'train' and [:13] for 'val'
        if i in (0, 13):
            vol = np.clip(np.round(rng.normal(0, 300, size=(40, 40, 24))), -1024, 3071).astype(np.int16)
            lab = np.zeros((40, 40, 24), dtype=np.int16)
            lab[10:30, 12:32, 6:18] = rng.integers(1, 8, size=(20, 20, 12))
            nifti.save(nifti.Nifti1Image(vol, np.diag([0.5, 0.5, 0.8, 1.0])), str(tmp_path / ("img%d.nii.gz" % i)))
            nifti.save(nifti.Nifti1Image(lab, np.diag([0.5, 0.5, 0.8, 1.0])), str(tmp_path / ("lab%d.nii.gz" % i)))
        j = 0 if i < 13 else 13
        entries.append({"image": str(tmp_path / ("img%d.nii.gz" % j)), "label": str(tmp_path / ("lab%d.nii.gz" % j))})
    (tmp_path / "dataset.json").write_text(json.dumps({"train_and_test": entries}))
    hm.args = types.SimpleNamespace(data=str(tmp_path) + "/")

    class SmallConfig(hm.HeartConfig):
        IMAGE_MIN_DIM = 64
        IMAGE_MAX_DIM = 64
        RPN_ANCHOR_SCALES = (16, 32)

    cfg = SmallConfig("beginning")
    ds = hm.HeartDataset()
    ds.load_heart("train")
    ds.prepare()
    assert ds.num_classes == 8 and len(ds.image_ids) == 1
    item = M.Dataset(ds, cfg)[0]
    image, meta, mask = item
    assert image.shape == (64, 64, 64, 1) and mask.shape == (64, 64, 64)
    assert set(np.unique(mask)) <= set(range(8)) and mask.max() > 0
    anchors = U.generate_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, cfg.RPN_ANCHOR_RATIOS, M.compute_backbone_shapes(cfg, cfg.IMAGE_SHAPE),
                                         cfg.BACKBONE_STRIDES, cfg.RPN_ANCHOR_STRIDE)
    np.random.seed(0)
    images, rpn_match, rpn_bbox, class_ids, boxes, masks = M.load_image_gt(image, mask, 0, ds, cfg, anchors)
    assert images.shape == (1, 64, 64, 64) and images.dtype == np.float32 and abs(float(images.mean())) < 1e-3
    assert rpn_match.shape == (anchors.shape[0], 1) and (rpn_match == 1).sum() > 0
    assert class_ids.tolist() == list(range(1, 8)) and boxes.shape == (7, 6) and masks.shape == (8, 64, 64, 64)
    z1, y1, x1, z2, y2, x2 = boxes[0]
    assert 0 <= z1 < z2 <= 64 and 0 <= y1 < y2 <= 64 and 0 <= x1 < x2 <= 64


def test_train_bn_is_refused_not_ignored():
    """BatchNorm runs frozen on this path (reference model.py:1297-1304,1401-1406); TRAIN_BN = True has no implementation
    and must raise instead of silently training nothing."""
    from cfun_b200 import config as Cf
    src = open(os.path.join(ROOT, "cfun_b200", "model.py")).read()
    assert 'raise NotImplementedError("config.TRAIN_BN = True is not supported' in src
    cfg = Cf.heart_config(64, "beginning", mask_pool=32, anchor_scales=(16, 32), TRAIN_BN=True)
    assert cfg.TRAIN_BN is True


def test_reference_arm_does_not_load_the_native_library():
    """`bench.py --impl reference` must not map libcfun_b200.so into its process: its inputs and weights come
    from cfun_b200.workload / cfun_b200.config, which import neither ops nor model."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.argv=['bench.py']; import bench; import cfun_b200.workload, cfun_b200.config; "
            "maps = open('/proc/self/maps').read(); assert 'libcfun_b200' not in maps, 'native library mapped'; "
            "assert 'cfun_b200.model' not in sys.modules and 'cfun_b200.ops' not in sys.modules; print('clean')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-2000:]


def test_static_rpn_losses_equal_the_reference_shaped_losses_on_cpu():
    """model.compute_rpn_losses_static (fixed-size gathers, no read-back; pure torch index algebra, so it runs here) against
    compute_rpn_class_loss / compute_rpn_bbox_loss (reference model.py:836-873) and against the oracle, values and gradients;
    also the scatter fallback of the padded nonzero"""
    from cfun_b200 import model as M
    g = torch.Generator().manual_seed(3)
    A, K = 3000, 64
    for npos, nneg in ((17, 40), (32, 32), (1, 0), (0, 5)):
        m = torch.zeros(A, dtype=torch.int32)
        perm = torch.randperm(A, generator=g)
        m[perm[:npos]] = 1
        m[perm[npos:npos + nneg]] = -1
        tgt = torch.zeros(1, K, 6)
        tgt[0, :npos] = torch.randn(npos, 6, generator=g)
        logits = torch.randn(1, A, 2, generator=g)
        pred = torch.randn(1, A, 6, generator=g) * 2
        mm = m.view(1, -1, 1)
        l0, p0 = logits.clone().requires_grad_(True), pred.clone().requires_grad_(True)
        c0, b0 = M.compute_rpn_class_loss(mm, l0), M.compute_rpn_bbox_loss(tgt, mm, p0)
        l1, p1 = logits.clone().requires_grad_(True), pred.clone().requires_grad_(True)
        c1, b1, counts = M.compute_rpn_losses_static(mm, tgt, l1, p1, K)
        assert counts.tolist() == [npos + nneg, npos]
        assert torch.allclose(c0, c1, rtol=1e-6, atol=0)
        assert (torch.isnan(b0).all() and torch.isnan(b1).all()) if npos == 0 else torch.allclose(b0, b1, rtol=1e-6, atol=0)
        (c0 + (b0 if npos else 0)).sum().backward()
        (c1 + (b1 if npos else 0)).sum().backward()
        assert torch.allclose(l0.grad, l1.grad, rtol=1e-5, atol=1e-9)
        if npos:
            assert torch.allclose(p0.grad, p1.grad, rtol=1e-5, atol=1e-9)
        if npos:
            assert abs(float(O.rpn_class_loss(m, logits[0])) - float(c1)) < 1e-6 and abs(float(O.rpn_bbox_loss(tgt[0], m, pred[0])) - float(b1)) < 1e-6
    mask = torch.rand(500, generator=g) > 0.9
    want = torch.nonzero(mask)[:, 0]
    for size in (int(mask.sum()) + 7, int(mask.sum())):
        got = M._nonzero_static(mask, size)
        real = torch.nonzero_static
        try:
            del torch.nonzero_static          # force the cumsum / scatter fallback
            fb = M._nonzero_static(mask, size)
        finally:
            torch.nonzero_static = real
        for idx in (got, fb):
            assert idx.shape[0] == size and torch.equal(idx[:want.numel()], want) and bool((idx[want.numel():] == -1).all())
