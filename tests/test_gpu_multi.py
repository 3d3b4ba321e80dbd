"""Real-NCCL checks of the data-parallel path: run with `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`.
On a single-GPU box these tests skip."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from conftest import load_golden
    from detweights import det_state
    from synth import golden_step_inputs
    from cfun_b200 import config as Cf, model as M
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    g = load_golden("step64_beginning")
    cfg = Cf.heart_config(64, "beginning", mask_pool=32, anchor_scales=(16, 32))
    net = M.MaskRCNN(cfg, "/tmp/_cfun_test")
    net.load_state_dict(det_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=int(g["seed_weights"])), strict=True)
    net = net.to(dev)
    inp = golden_step_inputs(g)
    net.mask.modified_u_net.injected_drop = inp["drop"]
    opt = net.make_optimizer(cfg.LEARNING_RATE)
    if overlap:
        opt.overlap_allreduce()

    def grads_for(volume_rank):
        # every rank owns its own volume: the golden volume, flipped along D for rank 1 (same label geometry by symmetry)
        image = inp["image"].to(dev)
        masks = inp["gt_masks"].to(dev)
        if volume_rank == 1:
            image = torch.flip(image, (3,)).contiguous()
            masks = torch.flip(masks, (2,)).contiguous()
        opt.zero_grad()
        torch.manual_seed(int(g["seed_perm"]))
        net.forward_backward(image, None, inp["rpn_match"].to(dev)[None, :, None], inp["rpn_bbox"].to(dev)[None],
                             torch.arange(1, 8).int().to(dev)[None], inp["gt_boxes"].to(dev)[None], masks[None])
        torch.cuda.synchronize()

    # the data-parallel step's gradient: own volume, then the ONE collective
    grads_for(rank)
    opt.allreduce_grads()
    torch.cuda.synchronize()
    reduced = opt.flat_grad.detach().clone()
    if rank == 0:
        # sum of the two single-GPU gradients, computed without any collective (hooks idle: pending list is cleared by zero_grad)
        saved_dist = opt._distributed
        opt._distributed = lambda: False
        total = torch.zeros_like(reduced)
        for r in range(world):
            grads_for(r)
            total += opt.flat_grad
        opt._distributed = saved_dist
        np.save(os.path.join(out_dir, "reduced_overlap%d.npy" % int(overlap)), reduced.cpu().numpy())
        np.save(os.path.join(out_dir, "sum_overlap%d.npy" % int(overlap)), total.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_nccl_allreduce_equals_sum_of_single_gpu_gradients(tmp_path, overlap):
    """FlatSGD.flat_grad after allreduce_grads() == sum over ranks of the single-GPU gradients of their volumes (<= 1e-6 of
    the gradient's max magnitude), with the plain single collective and with the overlapped early launch of the
    classifier.conv1 slice from its gradient hook."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / ("reduced_overlap%d.npy" % int(overlap)))
    tot = np.load(tmp_path / ("sum_overlap%d.npy" % int(overlap)))
    assert red.shape == tot.shape and np.isfinite(red).all()
    assert np.abs(red - tot).max() <= 1e-6 * np.abs(tot).max(), (np.abs(red - tot).max(), np.abs(tot).max())
    assert np.abs(tot).max() > 0
