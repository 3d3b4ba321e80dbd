"""GPU parity tests for the individual CUDA ops (through the C ABI) against the CPU oracle / torch-CPU fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cfun_oracle as O
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4          # north_star: fp32 conv activations within 1e-4 relative (max|diff| / max|ref|)


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from cfun_b200 import ops as _ops
    return _ops


def cuda(t):
    return t.cuda()


CONV_CASES = [
    # N, Cin, D, H, W, Cout, k, stride, pad, bias
    (1, 1, 16, 18, 20, 16, (3, 7, 7), 2, (1, 3, 3), True),      # stem (backbone.py:124)
    (1, 16, 8, 8, 8, 16, (1, 3, 3), 1, (0, 1, 1), True),        # conv_S
    (1, 16, 8, 8, 8, 16, (3, 1, 1), 1, (1, 0, 0), True),        # conv_T
    (1, 16, 8, 10, 12, 64, 1, 2, 0, True),                      # strided 1x1x1 downsample
    (2, 20, 9, 10, 11, 20, 3, 1, 1, False),                     # U-Net thin conv, ragged extents
    (2, 20, 12, 12, 12, 40, 3, 2, 1, False),                    # U-Net stride-2
    (1, 8, 10, 10, 10, 8, 5, 1, 2, False),                      # out_upscale_conv 5^3
    (1, 32, 6, 6, 6, 2, 1, 1, 0, True),                         # RPN class head
    (1, 32, 6, 6, 6, 6, 1, 1, 0, True),                         # RPN bbox head
    (3, 40, 9, 10, 11, 8, 1, 1, 0, False),                      # U-Net segmentation head conv3d_l4 (40 -> 8): pw_wgrad2_kernel, ragged chunk
    (2, 160, 5, 6, 7, 8, 1, 1, 0, False),                       # ds2_1x1_conv3d (160 -> 8): 160 thread tiles, one voxel group
    (1, 256, 8, 8, 8, 2, 1, 1, 0, True),                        # RPN class head at full width (256 -> 2)
    (1, 12, 5, 5, 5, 5, 1, 1, 0, False),                        # odd Cout (padded to 6 in the co tile)
    (2, 40, 21, 30, 33, 8, 1, 1, 0, True),                      # streaming pointwise kernels (pw_fwd2 / pw_dgrad2): several tiles per block, ragged last tile
    (1, 80, 17, 20, 24, 8, 1, 1, 0, False),                     # ds3_1x1_conv3d (80 -> 8): 128-voxel tiles
    (1, 32, 16, 18, 20, 6, 1, 1, 0, True),                      # RPN bbox head, Cout not a multiple of 4 (plain dY staging in the weight gradient)
    (1, 1, 12, 12, 12, 20, 3, 1, 1, False),                     # U-Net first conv (Cin = 1)
    (1, 48, 8, 8, 8, 80, 3, 1, 1, True),                        # >64 output channels
    (3, 24, 5, 7, 6, 36, 3, 1, 1, True),                        # odd everything
    (1, 1, 12, 20, 70, 24, (5, 7, 7), 2, (2, 3, 3), True),      # LiTS stem (LiTS_2017/backbone.py:124): conv_c1.cu <24;5,7,7;2>
    (2, 1, 9, 10, 37, 32, 3, 1, 1, False),                      # LiTS U-Net first conv (base 32): conv_c1.cu <32;3,3,3;1>
    (1, 1, 21, 9, 66, 16, (3, 7, 7), 2, (1, 3, 3), True),       # heart stem, ragged tiles in every direction
    (4, 1, 6, 7, 34, 20, 3, 1, 1, False),                       # heart U-Net first conv, ragged
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3d_fwd_bwd(ops, case):
    N, Cin, D, H, W, Cout, k, s, p, bias = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(N, Cin, D, H, W, generator=g)
    kk = (k, k, k) if isinstance(k, int) else k
    w = torch.randn(Cout, Cin, *kk, generator=g) * 0.2
    b = torch.randn(Cout, generator=g) if bias else None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    yr = F.conv3d(xr, wr, br, stride=s, padding=p)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xc, wc = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True)
    bc = cuda(b).requires_grad_(True) if bias else None
    yc = ops.conv3d(xc, wc, bc, s, p)
    assert tuple(yc.shape) == tuple(yr.shape)
    assert ops.is_cl(yc)
    yc.backward(cuda(dy))
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL
    if bias:
        assert rel_err(bc.grad.cpu().numpy(), br.grad.numpy()) < TOL


def test_conv3d_relu_epilogue(ops):
    g = torch.Generator().manual_seed(3)
    x, w, b = torch.randn(1, 16, 6, 6, 6, generator=g), torch.randn(24, 16, 3, 3, 3, generator=g) * 0.1, torch.randn(24, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.relu(F.conv3d(xr, wr, b, padding=1))
    yr.square().sum().backward()
    xc, wc = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True)
    yc = ops.conv3d(xc, wc, cuda(b), 1, 1, relu=True)
    yc.square().sum().backward()
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL


@pytest.mark.parametrize("M,Nout,C,pool", [(12, 128, 8, 6), (5, 16, 32, 4), (40, 24, 4, 5), (200, 16, 4, 4)])
def test_fc_conv(ops, M, Nout, C, pool):
    """any number of RoI rows: 40 = three 16-row chunks in backward, 200 (base Config TRAIN_ROIS_PER_IMAGE) = two 128-row
    chunks forward and thirteen backward"""
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, C, pool, pool, pool, generator=g)
    w = torch.randn(Nout, C, pool, pool, pool, generator=g) * 0.05
    b = torch.randn(Nout, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br)
    xc, wc, bc = cuda(x).requires_grad_(True), cuda(w).requires_grad_(True), cuda(b).requires_grad_(True)
    yc = ops.fc_conv(xc, wc, bc)
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    yc.backward(cuda(dy))
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL
    assert rel_err(bc.grad.cpu().numpy(), br.grad.numpy()) < TOL


@pytest.mark.parametrize("C,up,use_drop,dims", [(20, 1, False, (5, 6, 7)), (40, 2, False, (5, 6, 7)), (8, 1, True, (5, 6, 7)),
                                                (160, 1, True, (5, 6, 7)), (320, 2, False, (5, 6, 7)),
                                                # enough rows for several blocks and the unrolled main loops
                                                (20, 1, True, (40, 36, 44)), (40, 1, False, (33, 31, 29)),
                                                (8, 1, False, (48, 40, 40)), (6, 1, False, (17, 19, 23)),
                                                (16, 2, False, (20, 24, 28))])
def test_instnorm_lrelu(ops, C, up, use_drop, dims):
    g = torch.Generator().manual_seed(C + up)
    N, (D, H, W) = 2, dims
    x = torch.randn(N, C, D, H, W, generator=g) * 2 + 0.5
    drop = ((torch.rand(N, C, generator=g) > 0.6).float() / 0.4) if use_drop else None
    xr = x.clone().requires_grad_(True)
    t = xr * drop.view(N, C, 1, 1, 1) if use_drop else xr
    yr = F.leaky_relu(F.instance_norm(t, eps=1e-5), 0.01)
    if up == 2:
        yr = F.interpolate(yr, scale_factor=2, mode="nearest")
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xc = cuda(x).requires_grad_(True)
    yc = ops.instnorm_lrelu(xc, cuda(drop) if use_drop else None, 1e-5, 0.01, up)
    yc.backward(cuda(dy))
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < 2 * TOL


@pytest.mark.parametrize("C,G,P,dims", [(20, 3, 1, (6, 10, 12)), (40, 5, 1, (9, 16, 24)), (40, 5, 2, (5, 8, 8)), (16, 2, 1, (7, 9, 11)),
                                        (20, 4, 1, (3, 5, 7)), (80, 10, 1, (4, 12, 12)), (320, 40, 1, (3, 6, 6)), (6, 1, 2, (4, 5, 6))])
def test_pack_act_gp_layout(ops, C, G, P, dims):
    """the operand pack of the tcgen05 convolutions: group-planar [G][N*(D+2P)][H][W][8] split bf16 with P zero planes on
    both sides of every sample; hi = bf16(x), lo = bf16(x - hi), channels >= C zero -- bit-exact"""
    from cfun_b200.ops import _run, _ptr, _stream
    g = torch.Generator().manual_seed(C * 7 + P)
    N, (D, H, W) = 2, dims
    x = torch.randn(N, D, H, W, C, generator=g) * 3
    xc = x.cuda()
    hi = torch.full((G, N * (D + 2 * P), H, W, 8), 7.0, dtype=torch.bfloat16, device="cuda")
    lo = torch.full_like(hi, 7.0)
    _run("cfun_pack_act_gp", _ptr(xc), _ptr(hi), _ptr(lo), N, D, H, W, C, G, P, _stream())
    xp = torch.zeros(N, D + 2 * P, H, W, G * 8)
    xp[:, P:P + D, :, :, :C] = x
    eh = xp.to(torch.bfloat16)
    el = (xp - eh.float()).to(torch.bfloat16)
    to_gp = lambda t: t.reshape(N * (D + 2 * P), H, W, G, 8).permute(3, 0, 1, 2, 4).contiguous()
    assert torch.equal(hi.cpu().view(torch.int16), to_gp(eh).view(torch.int16))
    assert torch.equal(lo.cpu().view(torch.int16), to_gp(el).view(torch.int16))


def test_cat_channels_fwd_bwd(ops):
    """ops.cat_channels == torch.cat(dim=1) on channels-last tensors, and its backward splits the gradient"""
    g = torch.Generator().manual_seed(5)
    a = torch.randn(2, 20, 5, 6, 7, generator=g)
    b = torch.randn(2, 40, 5, 6, 7, generator=g)
    ac, bc = ops.to_cl(a.cuda()).requires_grad_(True), ops.to_cl(b.cuda()).requires_grad_(True)
    out = ops.cat_channels(ac, bc)
    assert ops.is_cl(out) and torch.equal(out.cpu(), torch.cat([a, b], 1))
    w = torch.randn(out.shape, generator=g)
    (out * w.cuda()).sum().backward()
    assert torch.equal(ac.grad.cpu(), w[:, :20]) and torch.equal(bc.grad.cpu(), w[:, 20:])
    odd = ops.cat_channels(ac[:, :3], bc)          # channel counts that are not multiples of 4 fall back to torch.cat
    assert torch.equal(odd.detach().cpu(), torch.cat([a[:, :3], b], 1))


def test_affine_act_residual_and_pool(ops):
    g = torch.Generator().manual_seed(5)
    x, r = torch.randn(2, 16, 4, 6, 8, generator=g), torch.randn(2, 16, 4, 6, 8, generator=g)
    a, b = torch.rand(16, generator=g) + 0.5, torch.randn(16, generator=g)
    xr, rr = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    yr = F.max_pool3d(F.relu(xr * a.view(1, -1, 1, 1, 1) + b.view(1, -1, 1, 1, 1) + rr), 2, 2)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xc, rc = cuda(x).requires_grad_(True), cuda(r).requires_grad_(True)
    yc = ops.maxpool2(ops.affine_act(xc, cuda(a), cuda(b), rc, 0.0, 1))
    yc.backward(cuda(dy))
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < 1e-6
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < 1e-6
    assert rel_err(rc.grad.cpu().numpy(), rr.grad.numpy()) < 1e-6
    # plain leaky relu and upsample
    xc2 = cuda(x).requires_grad_(True)
    y2 = ops.upsample2x(ops.leaky_relu(xc2))
    xr2 = x.clone().requires_grad_(True)
    y2r = F.interpolate(F.leaky_relu(xr2, 0.01), scale_factor=2, mode="nearest")
    dy2 = torch.randn(y2r.shape, generator=g)
    y2r.backward(dy2); y2.backward(cuda(dy2))
    assert rel_err(y2.detach().cpu().numpy(), y2r.detach().numpy()) < 1e-6
    assert rel_err(xc2.grad.cpu().numpy(), xr2.grad.numpy()) < 1e-5


def test_roi_crop_resize_matches_golden_and_grads(ops):
    gd = load_golden("roialign")
    f2, f3, boxes = torch.from_numpy(gd["f2"]), torch.from_numpy(gd["f3"]), torch.from_numpy(gd["boxes"])
    pool = tuple(int(p) for p in gd["pool"])
    lv = ops.roi_level(cuda(boxes))
    assert np.array_equal(lv.cpu().numpy(), gd["level"] - 2)
    single = ops.roi_crop_resize(cuda(f2)[None], None, cuda(boxes), None, pool, True)
    assert rel_err(single.cpu().numpy(), gd["single_level"]) < 1e-5
    f2c, f3c = cuda(f2)[None].requires_grad_(True), cuda(f3)[None].requires_grad_(True)
    pooled = ops.roi_crop_resize(f2c, f3c, cuda(boxes), lv, pool, False)
    assert rel_err(pooled.detach().cpu().numpy(), gd["pooled"]) < 1e-5
    assert float(pooled[0].abs().max()) == 0.0
    f2r, f3r = f2.clone().requires_grad_(True), f3.clone().requires_grad_(True)
    pr = O.pyramid_roi_align(boxes, [f2r, f3r], pool)
    w = torch.randn(pr.shape, generator=torch.Generator().manual_seed(1))
    (pr * w).sum().backward()
    (pooled * cuda(w)).sum().backward()
    assert rel_err(f2c.grad[0].cpu().numpy(), f2r.grad.numpy()) < 1e-5
    assert rel_err(f3c.grad[0].cpu().numpy(), f3r.grad.numpy()) < 1e-5


def test_sort_desc_total_order(ops):
    g = torch.Generator().manual_seed(2)
    for n in (1, 2, 37, 1000, 2048, 2049, 36864, 70000):
        s = torch.randn(n, generator=g)
        s[::7] = s[0]           # ties
        order = ops.sort_desc(cuda(s)).cpu().numpy()
        assert np.array_equal(order, O.sort_desc(s.numpy())), n


@pytest.mark.parametrize("case", ["rand_t7_m50", "rand_t3_all", "rand_t5_m1", "nested_degenerate", "integer_boxes_t3"])
def test_nms_bit_exact_golden(ops, case):
    from cfun_b200 import utils as U
    g = load_golden("nms")
    keep = U.non_max_suppression(g[case + "/boxes"], g[case + "/scores"], float(g[case + "/thr"]), int(g[case + "/max"]))
    assert keep.dtype == np.int32 and np.array_equal(keep, g[case + "/keep"])
    iou = U.compute_iou(g[case + "/boxes"][0], g[case + "/boxes"], None, None)
    assert np.array_equal(iou.view(np.uint32), g[case + "/iou0"].view(np.uint32))


@pytest.mark.parametrize("n,thr,mx", [(3000, 0.7, 500), (3000, 0.3, 3000), (257, 0.5, 64)])
def test_nms_bit_exact_random(ops, n, thr, mx):
    from cfun_b200 import utils as U
    rng = np.random.default_rng(n)
    c = rng.uniform(0, 256, size=(n, 3)); s = rng.uniform(16, 128, size=(n, 3))
    b = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.float32)
    sc = rng.uniform(0, 1, size=n).astype(np.float32)
    assert np.array_equal(U.non_max_suppression(b, sc, thr, mx), O.non_max_suppression(b, sc, thr, mx))


def _config5_boxes(n=10000):
    """BASELINE.json config 5 / SURVEY.md 8d: 10 000 random 3-D proposals over a 256^3 map, default_rng(0)"""
    rng = np.random.default_rng(0)
    c = rng.uniform(0, 256, size=(n, 3)); s = rng.uniform(16, 128, size=(n, 3))
    b = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 256).astype(np.float32)
    sc = rng.uniform(0, 1, size=n).astype(np.float32)
    return b, sc


@pytest.mark.parametrize("thr,mx", [(0.7, 500), (0.7, 10000), (0.3, 10000)])
def test_nms_bit_exact_config5_10k(ops, thr, mx):
    """full-size micro-benchmark configuration: the kept indices equal the numpy restatement, in order"""
    from cfun_b200 import utils as U
    b, sc = _config5_boxes()
    got = U.non_max_suppression(b, sc, thr, mx)
    want = O.non_max_suppression(b, sc, thr, mx)
    assert got.dtype == np.int32 and np.array_equal(got, want), (len(got), len(want))
    # size-independent properties: kept boxes are mutually below the threshold, in descending score order
    assert np.all(np.diff(sc[got]) <= 0)
    vol = ((b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1])) * (b[:, 5] - b[:, 2])
    k = got[:200]
    for i, a in enumerate(k[:-1]):
        assert np.all(O.compute_iou(b[a], b[k[i + 1:]], vol[a], vol[k[i + 1:]]) <= np.float32(thr))


def test_roi_crop_resize_config5_10k(ops):
    """10 000 boxes over a [1,256,256,256] map (the mask-branch case, model.py:1413), pool 12^3: every 25th box against the
    CPU restatement; all-zero rows exactly where the oracle leaves zeros"""
    b, _ = _config5_boxes()
    boxes = torch.from_numpy(b / 256.0)
    g = torch.Generator().manual_seed(11)
    fmap = torch.randn(1, 256, 256, 256, generator=g)
    out = ops.roi_crop_resize(cuda(fmap)[None], None, cuda(boxes), None, (12, 12, 12), True)
    assert tuple(out.shape) == (10000, 1, 12, 12, 12)
    sel = torch.arange(0, 10000, 25)
    ref = O.roi_align(fmap, (12, 12, 12), boxes[sel])
    assert rel_err(out[sel.cuda()].cpu().numpy(), ref.numpy()) < 1e-5


def test_decode_clip_and_proposals(ops):
    from cfun_b200 import model as M
    g = load_golden("proposal")
    anchors, probs, deltas = cuda(torch.from_numpy(g["anchors"])), cuda(torch.from_numpy(g["probs"])), cuda(torch.from_numpy(g["deltas"]))
    dec = M.apply_box_deltas(anchors, deltas * 0.1)
    assert rel_err(dec.cpu().numpy(), g["decoded"]) < 1e-6       # expf vs torch-CPU exp may differ in the last ulp
    assert np.array_equal(M.clip_boxes(cuda(torch.from_numpy(g["decoded"])), [0, 0, 0, 64, 64, 64]).cpu().numpy(), g["clipped"])

    class Cfg: PRE_NMS_LIMIT = 1000; IMAGE_SHAPE = g["image_shape"]; RPN_BBOX_STD_DEV = np.array([.1, .1, .1, .2, .2, .2])
    for key, count in (("rois_training", 500), ("rois_inference", 64)):
        rois = M.proposal_layer([probs[None], deltas[None]], count, 0.7, anchors, Cfg)[0]
        assert rois.shape == g[key].shape
        assert rel_err(rois.cpu().numpy(), g[key]) < 1e-6


def test_overlaps_refinement_targets(ops):
    from cfun_b200 import model as M
    g = load_golden("dtl")
    props, gtb = cuda(torch.from_numpy(g["proposals"])), cuda(torch.from_numpy(g["gt_boxes"]))
    assert np.array_equal(M.bbox_overlaps(props, gtb).cpu().numpy(), g["overlaps"])
    ref = ops.box_refinement(props[:20], gtb[:1].repeat(20, 1))
    assert rel_err(ref.cpu().numpy(), g["refinement"]) < 1e-6
    lab = cuda(torch.from_numpy(g["label"]))

    class Cfg:
        DETECTION_TARGET_IOU_THRESHOLD = 0.5; TRAIN_ROIS_PER_IMAGE = 15; ROI_POSITIVE_RATIO = 0.33
        BBOX_STD_DEV = np.array([.1, .1, .1, .2, .2, .2]); MASK_SHAPE = tuple(int(m) for m in g["mask_shape"])
        DENSE_MASK_TARGETS = True
    for dense in (True, False):
        Cfg.DENSE_MASK_TARGETS = dense
        torch.manual_seed(int(g["seed"]))
        p_rois, rois, cls, dl, msk = M.detection_target_layer(props[None], torch.arange(1, 8).int().cuda()[None], gtb[None], lab, Cfg)
        assert np.array_equal(p_rois.cpu().numpy(), g["positive_rois"])
        assert np.array_equal(rois.cpu().numpy(), g["rois"])
        assert np.array_equal(cls.cpu().numpy(), g["class_ids"])
        assert rel_err(dl.cpu().numpy(), g["deltas"]) < 1e-6
        if dense:
            assert msk.dtype == torch.float64 and np.array_equal(msk.cpu().numpy().astype(np.uint8), g["masks"])
        else:
            assert msk.dtype == torch.int64 and np.array_equal(msk.cpu().numpy(), g["masks"].argmax(1))
    # the one-hot stack form of gt_masks (reference call shape) gives the same targets
    onehot = torch.stack([(lab == c) for c in range(8)]).float()
    torch.manual_seed(int(g["seed"]))
    out = M.detection_target_layer(props[None], torch.arange(1, 8).int().cuda()[None], gtb[None], onehot[None], Cfg)
    assert np.array_equal(out[4].cpu().numpy(), g["masks"].argmax(1))


def test_refine_detections(ops):
    from cfun_b200 import model as M
    g = load_golden("refine")

    class Cfg:
        RPN_BBOX_STD_DEV = np.array([.1, .1, .1, .2, .2, .2]); IMAGE_SHAPE = (64, 64, 64, 1)
        DETECTION_MIN_CONFIDENCE = 0.7; DETECTION_NMS_THRESHOLD = 0.3; DETECTION_MAX_INSTANCES = 32
    det = M.refine_detections(cuda(torch.from_numpy(g["rois"])), cuda(torch.from_numpy(g["probs"])),
                              cuda(torch.from_numpy(g["deltas"])), [0, 0, 0, 64, 64, 64], Cfg)
    assert det.shape == g["detections"].shape
    assert rel_err(det.cpu().numpy(), g["detections"]) < 1e-6


def test_sobel_edge_loss_and_other_losses(ops):
    from cfun_b200 import model as M
    g = load_golden("losses")
    lab = torch.from_numpy(g["target_label"])
    tcls = torch.from_numpy(g["target_class_ids"])
    mlog = cuda(torch.from_numpy(g["mask_logits"])).requires_grad_(True)
    mprob = torch.softmax(mlog, 1)
    mprob.retain_grad()
    l_mask = M.compute_mrcnn_mask_loss(cuda(lab), cuda(tcls), mlog)
    l_edge = M.compute_mrcnn_mask_edge_loss(cuda(lab), cuda(tcls), mprob)
    assert abs(float(l_mask) - float(g["mask_loss"])) < 1e-5 * abs(float(g["mask_loss"]))
    assert abs(float(l_edge) - float(g["edge_loss"])) < 1e-4 * abs(float(g["edge_loss"]))
    (g_edge,) = torch.autograd.grad(l_edge.sum(), mprob, retain_graph=True)
    assert rel_err(g_edge.cpu().numpy(), g["g_edge"]) < TOL
    (g_mask,) = torch.autograd.grad(l_mask, mlog)
    assert rel_err(g_mask.cpu().numpy(), g["g_mask"]) < TOL
    # one-hot float64 target form (the reference's) gives the same numbers
    onehot = torch.stack([(lab == c) for c in range(8)], 1).double().cuda()
    assert abs(float(M.compute_mrcnn_mask_edge_loss(onehot, cuda(tcls), mprob)) - float(g["edge_loss"])) < 1e-4 * abs(float(g["edge_loss"]))
    rmatch = cuda(torch.from_numpy(g["rpn_match"]))
    assert abs(float(M.compute_rpn_class_loss(rmatch, cuda(torch.from_numpy(g["rpn_logits"])))) - float(g["rpn_class_loss"])) < 1e-5
    assert abs(float(M.compute_rpn_bbox_loss(cuda(torch.from_numpy(g["rpn_target"])), rmatch, cuda(torch.from_numpy(g["rpn_bbox"])))) - float(g["rpn_bbox_loss"])) < 1e-5


def test_mold_volume_and_sgd(ops):
    rng = np.random.default_rng(0)
    vol = np.clip(np.round(rng.normal(0, 300, size=(20, 24, 28))), -1024, 3071).astype(np.int16)     # [H,W,D]
    out = ops.mold_volume_i16(torch.from_numpy(vol).cuda())
    ref = O.mold_image(vol.astype(np.float32)[..., None]).transpose((3, 2, 0, 1))[None]
    assert rel_err(out.cpu().numpy(), ref) < 1e-5
    # fused clip + SGD(momentum, selective weight decay) against torch.optim.SGD + clip_grad_norm_
    from cfun_b200.dp import FlatSGD
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    ref_net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    ref_net.load_state_dict(net.state_dict())
    net = net.cuda()
    opt = FlatSGD(net, lr=0.1, momentum=0.9, weight_decay=1e-2, clip_norm=0.5)
    ropt = torch.optim.SGD(ref_net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-2)
    x = torch.randn(5, 8)
    for it in range(3):
        opt.zero_grad(); ropt.zero_grad()
        (net(x.cuda()) ** 2).sum().backward()
        (ref_net(x) ** 2).sum().backward()
        torch.nn.utils.clip_grad_norm_(ref_net.parameters(), 0.5)
        opt.step(); ropt.step()
    for p, q in zip(net.parameters(), ref_net.parameters()):
        assert rel_err(p.detach().cpu().numpy(), q.detach().numpy()) < 1e-5


def test_product_path_is_cuda_only(ops):
    with pytest.raises(RuntimeError):
        ops.conv3d(torch.randn(1, 4, 4, 4, 4), torch.randn(4, 4, 3, 3, 3), None, 1, 1)


# ---------------------------------------------------------------------------------------------------------
# tcgen05 implicit-GEMM path (forced), split-bf16 x3 = parity grade; single pass = fast mode
# ---------------------------------------------------------------------------------------------------------
TC_CASES = [
    # N, Cin, D, H, W, Cout, k, pad, bias, relu
    (2, 20, 12, 12, 16, 20, 3, 1, False, False),       # U-Net thin conv: Cin 20 -> K padded to 32, N = 32
    (1, 40, 16, 16, 24, 40, 3, 1, False, False),
    (1, 80, 9, 10, 11, 40, 3, 1, False, False),        # ragged extents: overhanging boxes, TMA zero fill
    (1, 128, 16, 16, 16, 256, 3, 1, True, True),       # RPN.conv_shared shape family, fused bias + ReLU
    (1, 16, 12, 12, 12, 320, 3, 1, True, False),       # two N tiles of 160
    (1, 32, 10, 10, 16, 16, 5, 2, False, False),       # 5^3 kernel (out_upscale_conv family)
    (1, 48, 8, 8, 20, 24, 3, 0, True, False),          # no padding
    (2, 64, 4, 4, 4, 96, 3, 1, True, True),            # one output tile, K split over 6 CTAs + ordered reduce (bias + ReLU there)
    (4, 320, 6, 6, 6, 320, 3, 1, False, False),        # bottom of the U-Net: few tiles x 540 K chunks, split K
]


HALO_CASES = [
    (2, 20, 12, 20, 16, 20, 3, 1, False, False),       # thin U-Net conv: K 20 -> 32 (2 chunks), N = 32
    (1, 40, 5, 33, 24, 40, 3, 1, True, True),          # K 48 (3 chunks), ragged H (3 slabs, last overhangs), bias + ReLU
    (1, 64, 4, 16, 8, 24, 3, 1, False, False),         # K 64 (4 chunks), exactly one slab per plane
    (3, 16, 3, 9, 11, 64, 3, 1, True, False),          # box larger than the map, W not a multiple of 8
]


@pytest.mark.parametrize("halo", ["1", "0"])
@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3d_tcgen05_halo_kernel_vs_tap_reload_kernel(ops, case, halo, monkeypatch):
    """the halo-resident kernel (conv_tc_halo.cu) and the per-tap reload kernel (conv_tc.cu) both match fp32"""
    monkeypatch.setenv("CFUN_TC_HALO", halo)
    test_conv3d_tcgen05_fwd_dgrad(ops, case)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv3d_tcgen05_fwd_dgrad(ops, case):
    N, Cin, D, H, W, Cout, k, p, bias, relu = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, k, generator=g) * (1.0 / (Cin * k ** 3) ** 0.5)
    b = torch.randn(Cout, generator=g) if bias else None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, b, padding=p)
    dy = torch.randn(yr.shape, generator=g)
    if relu:
        # the ReLU mask is decided by pre-activations that differ by ~1e-5 between the two paths; keep the comparison
        # well-posed by zeroing the incoming gradient wherever the pre-activation is within 1e-3 of the kink
        dy = dy * (yr.detach().abs() > 1e-3).float()
        yr = F.relu(yr)
    yr.backward(dy)
    # the weight gradient runs on tensor cores where a kernel supports the shape; otherwise only fwd + dgrad do
    wgrad_tc = ops.conv3d_supported(x.shape, w.shape, 1, p, ops.PASS_BWD_WEIGHT, ops.ALGO_TC)
    ops.set_conv_algo(ops.ALGO_TC)
    try:
        xc = x.cuda().requires_grad_(True)
        wc = w.cuda().requires_grad_(wgrad_tc)
        yc = ops.conv3d(xc, wc, b.cuda() if bias else None, 1, p, relu=relu)
        yc.backward(dy.cuda())
        torch.cuda.synchronize()
    finally:
        ops.set_conv_algo(ops.ALGO_AUTO)
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    if wgrad_tc:
        assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL


NEW_TC_CASES = [
    # N, Cin, D, H, W, Cout, k, stride, pad, what it covers
    (1, 20, 16, 16, 16, 40, 3, 2, 1, "stride 2 by space-to-depth (conv_s2d.cu), fwd + dgrad"),
    (2, 40, 8, 12, 16, 80, 3, 2, 1, "stride 2, ragged half extents"),
    (4, 64, 6, 6, 6, 48, 3, 1, 1, "6^3 volume: batch-dim boxes (bn = 4) in conv_tc.cu, padded lines in the wgrad pack"),
    (3, 32, 4, 4, 4, 32, 3, 1, 1, "4^3 volume, N not a multiple of the box"),
    (1, 320, 12, 12, 12, 24, 3, 1, 1, "Cin 320: two channel slices in conv_tc_wgrad.cu, W = 12 padded to 16"),
    (2, 24, 5, 18, 10, 40, 3, 1, 1, "d-stacked wgrad (conv_tc_wgrad_ds.cu): 3 + 5 channel groups, ragged slabs"),
    (1, 80, 6, 16, 16, 80, 3, 1, 1, "d-stacked wgrad: two M tiles of 5 groups, two tap groups"),
    (2, 16, 2, 8, 8, 16, 3, 1, 1, "d-stacked wgrad: two channel groups, D smaller than the 3-plane stack"),
    (1, 20, 3, 24, 24, 20, 3, 1, 1, "kw-stacked wgrad: 3 groups x 3 kw = 72 columns padded to N = 80, 8-line tiles"),
    (1, 16, 3, 40, 16, 16, 3, 1, 1, "kw-stacked wgrad: 16-line tiles (3 stages fit), ragged H (40 = 2.5 tiles)"),
    (1, 128, 3, 16, 16, 48, 3, 1, 1, "kw-stacked wgrad: Cin 128 = 3 channel slices (6, 6, 4 groups: the last one zero-filled)"),
    (1, 40, 4, 11, 13, 40, 3, 1, 1, "kw-stacked wgrad: W, H not multiples of 8 (TMA zero fill on both sides of every copy)"),
    (2, 8, 6, 16, 24, 8, 5, 1, 2, "5^3 out_upscale_conv (8 -> 8): halo kernel with 125 taps = 63 K-steps, 5-plane-stacked wgrad with 5 kw copies"),
    (1, 8, 3, 9, 13, 8, 5, 1, 2, "5^3, D smaller than the kernel, ragged H / W"),
    (1, 3, 4, 12, 20, 3, 5, 1, 2, "5^3 LiTS out_upscale_conv (3 -> 3): scalar pack path, weight gradient on CUDA cores"),
]


@pytest.mark.parametrize("case", NEW_TC_CASES, ids=[c[-1].split(":")[0].split(",")[0] + "_%d" % i for i, c in enumerate(NEW_TC_CASES)])
def test_conv3d_tcgen05_strided_tiny_and_stacked_wgrad(ops, case):
    N, Cin, D, H, W, Cout, k, st, p, _ = case
    g = torch.Generator().manual_seed(sum(case[:9]))
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, k, generator=g) * (1.0 / (Cin * k ** 3) ** 0.5)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, None, stride=st, padding=p)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    sup = [ops.conv3d_supported(x.shape, w.shape, st, p, ps, ops.ALGO_TC) for ps in (ops.PASS_FWD, ops.PASS_BWD_DATA, ops.PASS_BWD_WEIGHT)]
    assert sup[0] and sup[1], "the tensor-core path must take this shape"
    ops.set_conv_algo(ops.ALGO_TC)
    try:
        xc = x.cuda().requires_grad_(True)
        wc = w.cuda().requires_grad_(sup[2])
        yc = ops.conv3d(xc, wc, None, st, p)
        yc.backward(dy.cuda())
        torch.cuda.synchronize()
        assert ops.tc_debug_status() is None
    finally:
        ops.set_conv_algo(ops.ALGO_AUTO)
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    if st == 1 and (k == 3 or Cout % 4 == 0):
        assert sup[2], "stride-1 3^3 (and 8-channel 5^3) weight gradients run on tensor cores"
    if sup[2]:
        assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL


@pytest.mark.parametrize("case", [(2, 24, 12, 16, 16, 40, True), (1, 40, 12, 24, 16, 20, False), (1, 160, 5, 16, 16, 80, True),
                                  (1, 128, 8, 16, 16, 256, True), (2, 8, 6, 16, 24, 8, True, 5)])
def test_conv3d_fused_backward_keeps_forward_pack(ops, case):
    """AUTO picks the fused path (cfun_conv3d_fwd_keep_pack + cfun_conv3d_bwd_fused) for these shapes: same results as
    the three separate calls and as fp32, including the no-input-gradient case (first layer of a network)."""
    import ctypes as C
    from cfun_b200._lib import lib
    N, Cin, D, H, W, Cout, need_dx = case[:7]
    k = case[7] if len(case) > 7 else 3            # 5: out_upscale_conv (two zero planes per sample in the shared packs)
    g = torch.Generator().manual_seed(sum(case[:6]))
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, k, generator=g) * (1.0 / (Cin * k ** 3) ** 0.5)
    b = torch.randn(Cout, generator=g)
    xr, wr, br = x.clone().requires_grad_(need_dx), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, padding=k // 2)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    d = ops._conv_desc(x.shape, w.shape, 1, k // 2)
    assert lib.cfun_conv3d_pack_bytes(C.byref(d)) > 0, "the fused path must take this shape"
    xc, wc, bc = x.cuda().requires_grad_(need_dx), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    yc = ops.conv3d(xc, wc, bc, 1, k // 2)
    assert yc.grad_fn is not None and yc.grad_fn.fused
    yc.backward(dy.cuda())
    torch.cuda.synchronize()
    assert ops.tc_debug_status() is None
    assert rel_err(yc.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(wc.grad.cpu().numpy(), wr.grad.numpy()) < TOL
    assert rel_err(bc.grad.cpu().numpy(), br.grad.numpy()) < TOL
    if need_dx:
        assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < TOL
    else:
        assert xc.grad is None


def test_conv3d_tcgen05_single_pass_is_fast_mode_only(ops):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 64, 12, 12, 12, generator=g)
    w = torch.randn(64, 64, 3, 3, 3, generator=g) * 0.03
    yr = F.conv3d(x, w, None, padding=1)
    ops.set_conv_algo(ops.ALGO_TC1)
    try:
        y1 = ops.conv3d(x.cuda(), w.cuda(), None, 1, 1)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_algo(ops.ALGO_AUTO)
    e1 = rel_err(y1.cpu().numpy(), yr.numpy())
    assert 1e-4 < e1 < 3e-2, e1      # one bf16 pass misses the 1e-4 parity bar (why the x3 split is the default)


def rel_err_elementwise(a, b, floor=1e-1):
    """max over the elements with |b| > floor * max|b| of |a - b| / |b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.abs(b) > floor * np.abs(b).max()
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max())


@pytest.mark.parametrize("case", [(4, 40, 96, 40, "U-Net conv_norm_lrelu_l4.0: largest FLOP share of the step"),
                                  (1, 128, 32, 256, "RPN.conv_shared: north-star headline conv"),
                                  (4, 20, 96, 20, "U-Net conv3d_c1_2 / lrelu_conv_c1: Cin 20 = 3 real channel groups")],
                         ids=["unet40", "rpn128_256", "unet20"])
def test_conv3d_benchmarked_shapes_fwd_dgrad_wgrad(ops, case):
    """The shapes every performance claim rests on, at FULL size, all three passes, against a float64 cuDNN convolution of
    the same operands (exact to ~1e-15): max-normalised error < 1e-4 (the north-star bar), element-wise relative error
    < 1e-3 on every element above 10 % of the maximum (weight gradients are sums over 3.5 M voxels: the absolute error is
    uniform, so small elements carry a larger relative one), and an rms error below 2e-5 of the tensor's rms -- what 16 mantissa
    bits per operand (split-bf16) give: measured 5e-6, against 6e-7 for an fp32 cuDNN convolution; a single
    bf16 or tf32 pass sits at 2e-3 / 5e-4."""
    N, Ci, S, Co, _ = case
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Ci + Co)
    x = torch.randn(N, Ci, S, S, S, generator=g).to(dev)
    w = (torch.randn(Co, Ci, 3, 3, 3, generator=g) * (1.0 / (Ci * 27) ** 0.5)).to(dev)
    dy = torch.randn(N, Co, S, S, S, generator=g).to(dev)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y64 = F.conv3d(x64, w64, None, 1, 1)
    y64.backward(dy.double())
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x32, w32 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y32 = F.conv3d(x32, w32, None, 1, 1)
        y32.backward(dy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    xc, wc = ops.to_cl(x).requires_grad_(True), w.clone().requires_grad_(True)
    yc = ops.conv3d(xc, wc, None, 1, 1)
    assert yc.grad_fn.fused, "the benchmarked shapes run on the fused tcgen05 path"
    yc.backward(ops.to_cl(dy))
    torch.cuda.synchronize()
    for name, got, ref64, ref32 in (("fwd", yc, y64, y32), ("dgrad", xc.grad, x64.grad, x32.grad), ("wgrad", wc.grad, w64.grad, w32.grad)):
        gnp, r64 = got.detach().cpu().numpy().astype(np.float64), ref64.detach().cpu().numpy()
        assert rel_err(gnp, r64) < TOL, (name, rel_err(gnp, r64))
        assert rel_err_elementwise(gnp, r64) < 1e-3, (name, rel_err_elementwise(gnp, r64))
        scale = float(np.sqrt(np.mean(r64 ** 2)))
        rms = float(np.sqrt(np.mean((gnp - r64) ** 2))) / scale
        rms32 = float(np.sqrt(np.mean((ref32.detach().cpu().numpy().astype(np.float64) - r64) ** 2))) / scale
        # forward / data gradient contract 27 * Cin <= 3456 products; the weight gradient 3.5 M voxels, accumulated in the fp32
        # TMEM accumulator over 24 k-voxel split-K chunks per CTA and combined with fp32 atomics: measured 8e-5
        assert rms < (2e-4 if name == "wgrad" else 2e-5) and rms32 < rms, (name, rms, rms32)


@pytest.mark.parametrize("case", [(2, 8, 6, 7, 9, False), (3, 3, 5, 8, 6, True), (1, 5, 4, 4, 4, True)])
def test_mask_cross_entropy_fused(ops, case):
    """cfun_mask_ce_fwd / bwd against F.cross_entropy (mean reduction, optional class weights as in LiTS_2017/model.py:926)"""
    P, C, D, H, W, weighted = case
    g = torch.Generator().manual_seed(P * 100 + C)
    x = torch.randn(P, C, D, H, W, generator=g) * 3
    y = torch.randint(0, C, (P, D, H, W), generator=g)
    w = (torch.rand(C, generator=g) * 5 + 0.1) if weighted else None
    xr = x.clone().requires_grad_(True)
    lr = F.cross_entropy(xr, y, weight=w)
    (lr * 1.7).backward()
    xc = cuda(x).requires_grad_(True)
    lc = ops.mask_cross_entropy(xc, cuda(y), cuda(w) if weighted else None)
    (lc * 1.7).backward()
    assert abs(float(lc) - float(lr)) < 1e-5 * abs(float(lr))
    assert rel_err(xc.grad.cpu().numpy(), xr.grad.numpy()) < 1e-5


@pytest.mark.parametrize("case", [dict(N=2, Cin=20, Cout=20, dims=(6, 19, 27)), dict(N=3, Cin=40, Cout=24, dims=(5, 16, 16)),
                                  dict(N=4, Cin=40, Cout=40, dims=(20, 40, 48)), dict(N=2, Cin=24, Cout=72, dims=(5, 16, 16)),
                                  dict(N=2, Cin=80, Cout=80, dims=(4, 12, 20)), dict(N=1, Cin=128, Cout=256, dims=(4, 16, 16))])
def test_conv3d_epilogue_instnorm_statistics(ops, case):
    """ops.conv3d(in_stats=True): the per-(sample, channel) sums the conv epilogue accumulates equal the separate
    statistics pass over its output (cfun_instnorm_stats), so instnorm_lrelu gives the same result either way."""
    g = torch.Generator().manual_seed(case["Cin"])
    N, Cin, Cout, (D, H, W) = case["N"], case["Cin"], case["Cout"], case["dims"]
    x = cuda(torch.randn(N, Cin, D, H, W, generator=g)).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) * 0.1).cuda().requires_grad_(True)
    b = torch.randn(Cout, generator=g).cuda().requires_grad_(True)
    y = ops.conv3d(x, w, b, 1, 1, False, in_stats=True)
    acc = getattr(y, "_cfun_in_stats", None)
    assert acc is not None, "shape did not route to the halo-family kernels"
    yd = y.detach()
    ref = torch.stack([yd.double().sum(dim=(2, 3, 4)), yd.double().square().sum(dim=(2, 3, 4))], dim=2).reshape(-1)
    scale = torch.stack([yd.double().abs().sum(dim=(2, 3, 4)), yd.double().square().sum(dim=(2, 3, 4))], dim=2).reshape(-1)
    assert float(((acc - ref).abs() / scale).max()) < 2e-6          # fp32 partial sums, double totals
    out_fused = ops.instnorm_lrelu(y)
    y2 = yd.clone()
    out_sep = ops.instnorm_lrelu(y2)
    assert rel_err(out_fused.detach().cpu().numpy(), out_sep.cpu().numpy()) < 2e-6
    dy = cuda(torch.randn(out_fused.shape, generator=g))
    out_fused.backward(dy)          # gradients flow through norm and conv as before
    assert x.grad is not None and w.grad is not None and torch.isfinite(w.grad).all()


@pytest.mark.parametrize("case", [dict(N=2, Cin=20, Cout=20, dims=(6, 19, 27), drop=False, up=1), dict(N=3, Cin=40, Cout=20, dims=(5, 16, 16), drop=True, up=1),
                                  dict(N=2, Cin=80, Cout=80, dims=(4, 12, 20), drop=True, up=1), dict(N=2, Cin=40, Cout=40, dims=(7, 16, 24), drop=False, up=1),
                                  dict(N=1, Cin=160, Cout=80, dims=(24, 24, 24), drop=False, up=1)])
def test_conv_in_lrelu_fused_node_matches_separate_ops(ops, case):
    """ops.conv_in_lrelu (one autograd node: the norm backward writes the conv's output gradient straight into the split-bf16
    operand pack) == instnorm_lrelu(conv3d(x)): same forward, same input gradient bit for bit (the pack holds exactly
    split(dx)), weight gradient equal up to the atomics' summation order."""
    g = torch.Generator().manual_seed(case["Cin"] + case["Cout"])
    N, Cin, Cout, (D, H, W) = case["N"], case["Cin"], case["Cout"], case["dims"]
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * 0.1
    drop = ((torch.rand(N, Cout, generator=g) > 0.6).float() / 0.4).cuda() if case["drop"] else None
    # fp32 PyTorch reference first: the incoming gradient is zeroed wherever the normalised value is within 1e-3 of the
    # LeakyReLU kink, where paths that differ by ~1e-5 may pick different slopes
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    t = F.conv3d(xr, wr, None, padding=1)
    if case["drop"]:
        t = t * drop.cpu().view(N, Cout, 1, 1, 1)
    tn = F.instance_norm(t, eps=1e-5)
    zr = F.leaky_relu(tn, 0.01)
    dz = cuda(torch.randn(zr.shape, generator=g) * (tn.detach().abs() > 1e-3).float())
    zr.backward(dz.cpu())
    res = []
    for fused in (True, False):
        xc, wc = cuda(x).requires_grad_(True), w.cuda().requires_grad_(True)
        if fused:
            z = ops.conv_in_lrelu(xc, wc, None, 1, 1, drop)
            assert type(z.grad_fn).__name__.startswith("ConvInstNormActFn"), "shape did not take the fused node"
        else:
            z = ops.instnorm_lrelu(ops.conv3d(xc, wc, None, 1, 1), drop)
        z.backward(dz)
        res.append((z.detach(), xc.grad, wc.grad))
    (z1, dx1, dw1), (z0, dx0, dw0) = res
    assert rel_err(z1.cpu().numpy(), z0.cpu().numpy()) < 2e-6          # statistics: epilogue partial sums vs separate pass
    assert rel_err(dx1.cpu().numpy(), dx0.cpu().numpy()) < 1e-5
    assert rel_err(dw1.cpu().numpy(), dw0.cpu().numpy()) < 1e-5
    assert rel_err(z1.cpu().numpy(), zr.detach().numpy()) < TOL
    assert rel_err(dx1.cpu().numpy(), xr.grad.numpy()) < 5 * TOL


@pytest.mark.parametrize("C,dims", [(20, (9, 10, 11)), (20, (40, 36, 44)), (8, (12, 20, 28)), (6, (7, 9, 5))])
def test_add_lrelu_instnorm_fused_node(ops, C, dims):
    """ops.add_lrelu_instnorm(a, b) == (leaky_relu(a + b), leaky_relu(instance_norm(a + b))), values and both input gradients"""
    g = torch.Generator().manual_seed(C + dims[0])
    N, (D, H, W) = 2, dims
    a, b = torch.randn(N, C, D, H, W, generator=g) * 2, torch.randn(N, C, D, H, W, generator=g) + 0.3
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    sr = ar + br
    cr, yr = F.leaky_relu(sr, 0.01), F.leaky_relu(F.instance_norm(sr, eps=1e-5), 0.01)
    dc, dy = torch.randn(cr.shape, generator=g), torch.randn(yr.shape, generator=g)
    (cr * dc + yr * dy).sum().backward()
    ac, bc = cuda(a).requires_grad_(True), cuda(b).requires_grad_(True)
    c, y = ops.add_lrelu_instnorm(ac, bc)
    (c * cuda(dc) + y * cuda(dy)).sum().backward()
    assert rel_err(c.detach().cpu().numpy(), cr.detach().numpy()) < 1e-6
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) < TOL
    assert rel_err(ac.grad.cpu().numpy(), ar.grad.numpy()) < 2 * TOL
    assert torch.equal(ac.grad, bc.grad)


@pytest.mark.parametrize("case", [dict(N=2, C1=20, C2=20, Cout=40, dims=(6, 19, 27)), dict(N=3, C1=40, C2=40, Cout=80, dims=(5, 16, 16)),
                                  dict(N=1, C1=80, C2=80, Cout=160, dims=(6, 12, 12)), dict(N=2, C1=12, C2=28, Cout=24, dims=(4, 9, 17), fused=False)])     # last: a shape outside the fused backward -> cat + separate ops
def test_conv_in_lrelu_on_concatenated_inputs(ops, case):
    """ops.conv_in_lrelu(a, w, x2=b): the operand pack is built straight from the two tensors (cfun_conv3d_fwd_stats_cat) ==
    conv_in_lrelu(torch.cat((a, b), 1), w): same output, and the gradient of the concatenation split back to a and b"""
    g = torch.Generator().manual_seed(case["C1"] + case["Cout"])
    N, C1, C2, Cout, (D, H, W) = case["N"], case["C1"], case["C2"], case["Cout"], case["dims"]
    a, b = torch.randn(N, C1, D, H, W, generator=g), torch.randn(N, C2, D, H, W, generator=g)
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, generator=g) * 0.1
    res = []
    dz = None
    for fused in (True, False):
        ac, bc, wc = cuda(a).requires_grad_(True), cuda(b).requires_grad_(True), w.cuda().requires_grad_(True)
        if fused:
            z = ops.conv_in_lrelu(ac, wc, None, 1, 1, None, x2=bc)
            assert type(z.grad_fn).__name__.startswith("ConvInstNormActFn") == case.get("fused", True)
        else:
            z = ops.conv_in_lrelu(torch.cat((ac, bc), 1), wc, None, 1, 1, None)
        if dz is None:
            dz = cuda(torch.randn(z.shape, generator=g))
        z.backward(dz)
        res.append((z.detach(), ac.grad, bc.grad, wc.grad))
    (z1, da1, db1, dw1), (z0, da0, db0, dw0) = res
    assert torch.equal(z1, z0) or rel_err(z1.cpu().numpy(), z0.cpu().numpy()) < 2e-6
    assert rel_err(da1.cpu().numpy(), da0.cpu().numpy()) < 1e-5 and rel_err(db1.cpu().numpy(), db0.cpu().numpy()) < 1e-5
    assert rel_err(dw1.cpu().numpy(), dw0.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("case", [dict(N=2, Cin=20, Cout=20, dims=(6, 19, 27), drop=False), dict(N=3, Cin=20, Cout=20, dims=(5, 16, 16), drop=True),
                                  dict(N=2, Cin=40, Cout=24, dims=(7, 16, 24), drop=True)])
def test_lrelu_conv3d_activation_fused_into_the_operand_pack(ops, case):
    """ops.lrelu_conv3d: LeakyReLU (and the Dropout3d channel scale) applied on the way into the conv's operand pack ==
    conv3d(affine_act(x)): forward and input gradient bit for bit, weight gradient up to the atomics' summation order"""
    g = torch.Generator().manual_seed(case["Cin"] + case["dims"][0])
    N, Cin, Cout, (D, H, W) = case["N"], case["Cin"], case["Cout"], case["dims"]
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * 0.1
    scale = ((torch.rand(N, Cin, generator=g) > 0.6).float() / 0.4).cuda() if case["drop"] else None
    dy = None
    res = []
    for fused in (True, False):
        xc, wc = cuda(x).requires_grad_(True), w.cuda().requires_grad_(True)
        if fused:
            y = ops.lrelu_conv3d(xc, wc, None, 1, 1, scale)
            assert type(y.grad_fn).__name__.startswith("PreActConv3dFn"), "shape did not take the fused node"
        else:
            a = ops.leaky_relu(xc) if scale is None else ops.affine_act(xc, scale, torch.zeros_like(scale), None, 0.01, 1)
            y = ops.conv3d(a, wc, None, 1, 1)
        if dy is None:
            dy = cuda(torch.randn(y.shape, generator=g))
        y.backward(dy)
        res.append((y.detach(), xc.grad, wc.grad))
    (y1, dx1, dw1), (y0, dx0, dw0) = res
    assert torch.equal(y1, y0)
    assert torch.equal(dx1, dx0)
    assert rel_err(dw1.cpu().numpy(), dw0.cpu().numpy()) < 1e-5
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    t = xr if scale is None else xr * scale.cpu().view(N, Cin, 1, 1, 1)
    yr = F.conv3d(F.leaky_relu(t, 0.01), wr, None, padding=1)
    yr.backward(dy.cpu())
    assert rel_err(y1.cpu().numpy(), yr.detach().numpy()) < TOL and rel_err(dx1.cpu().numpy(), xr.grad.numpy()) < TOL


@pytest.mark.parametrize("case", [dict(N=2, C1=40, Cout=20, dims=(5, 8, 12)), dict(N=3, C1=80, Cout=40, dims=(4, 8, 8)),
                                  dict(N=1, C1=320, Cout=160, dims=(3, 6, 6)), dict(N=2, C1=20, Cout=20, dims=(3, 9, 7))])
def test_upnorm_conv_in_lrelu_fused_up_block(ops, case):
    """ops.upnorm_conv_in_lrelu (norm -> lrelu -> upsample x2 written straight into the conv's operand pack -> conv -> norm ->
    lrelu as one node) == the separate ops: forward and input gradient bit for bit, weight gradient up to the atomics'
    summation order; and against fp32 PyTorch"""
    g = torch.Generator().manual_seed(case["C1"] + case["dims"][1])
    N, C1, Cout, (D, H, W) = case["N"], case["C1"], case["Cout"], case["dims"]
    t = torch.randn(N, C1, D, H, W, generator=g) * 2 + 0.3
    w = torch.randn(Cout, C1, 3, 3, 3, generator=g) * 0.1
    # fp32 PyTorch reference; the incoming gradient is zeroed within 1e-3 of the second LeakyReLU's kink
    tr, wr = t.clone().requires_grad_(True), w.clone().requires_grad_(True)
    u = F.interpolate(F.leaky_relu(F.instance_norm(tr, eps=1e-5), 0.01), scale_factor=2, mode="nearest")
    yn = F.instance_norm(F.conv3d(u, wr, None, padding=1), eps=1e-5)
    zr = F.leaky_relu(yn, 0.01)
    dz = cuda(torch.randn(zr.shape, generator=g) * (yn.detach().abs() > 1e-3).float())
    zr.backward(dz.cpu())
    res = []
    for fused in (True, False):
        tc, wc = cuda(t).requires_grad_(True), w.cuda().requires_grad_(True)
        if fused:
            z = ops.upnorm_conv_in_lrelu(tc, wc)
            assert type(z.grad_fn).__name__.startswith("UpNormConvInstNormActFn"), "shape did not take the fused node"
        else:
            z = ops.conv_in_lrelu(ops.instnorm_lrelu(tc, None, 1e-5, 0.01, 2), wc, None, 1, 1)
        z.backward(dz)
        res.append((z.detach(), tc.grad, wc.grad))
    (z1, dt1, dw1), (z0, dt0, dw0) = res
    assert torch.equal(z1, z0)
    assert torch.equal(dt1, dt0)
    assert rel_err(dw1.cpu().numpy(), dw0.cpu().numpy()) < 1e-5
    assert rel_err(z1.cpu().numpy(), zr.detach().numpy()) < TOL
    assert rel_err(dt1.cpu().numpy(), tr.grad.numpy()) < 1e-3        # two InstanceNorm backward passes amplify the conv's 1e-5
