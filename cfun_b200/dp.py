"""Optimizer tail and volume-level data parallelism.

FlatSGD re-homes every trainable parameter, its gradient and its momentum in three flat fp32 buffers (parameters and
.grad become views), so that
  * the global-norm clip (clip_grad_norm_(params, 5.0), reference model.py:1641) + SGD(momentum, selective weight decay,
    reference model.py:1538-1545,1643) is one sum-of-squares kernel + one fused step kernel, and
  * data parallelism is exactly ONE NCCL all-reduce (SUM) of the flat gradient per optimizer step over NVLink/NVSwitch --
    the path shards on independent CT volumes; there is no other collective.
SUM (not mean) mirrors the reference's un-normalised gradient accumulation over BATCH_SIZE volumes (model.py:1640-1645).
"""
import torch
import torch.distributed as dist

from . import ops


class FlatSGD(object):
    def __init__(self, model, lr, momentum=0.9, weight_decay=1e-4, clip_norm=5.0, process_group=None):
        self.lr, self.momentum, self.weight_decay, self.clip_norm = lr, momentum, weight_decay, clip_norm
        self.group = process_group
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise RuntimeError("no trainable parameters")
        dev = named[0][1].device
        total = sum(p.numel() for _, p in named)
        self.flat_param = torch.empty(total, device=dev)
        self.flat_grad = torch.zeros(total, device=dev)
        self.flat_mom = torch.zeros(total, device=dev)
        self.wd_mask = torch.zeros(total, dtype=torch.uint8, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.names, self.slices = [], {}
        off = 0
        with torch.no_grad():
            for n, p in named:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                if 'bn' not in n:                       # reference: weight decay skips names containing 'bn'
                    self.wd_mask[off:off + k] = 1
                self.names.append(n)
                self.slices[n] = (off, off + k)
                off += k
        self.numel = total
        self._early = []          # (start, end) slices whose all-reduce is launched from a gradient hook
        self._pending = []        # (start, end, work) of this step
        self._side = None
        self._param_of = dict(named)

    def zero_grad(self):
        self.flat_grad.zero_()
        self._pending = []

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def overlap_allreduce(self, names=("classifier.conv1.weight",)):
        """Start the all-reduce of the named parameters' gradient slices as soon as autograd has produced them, on a side
        stream, so that it overlaps the rest of backward.  classifier.conv1.weight is 28.3 M of the 41.35 M parameters
        (113 of the 165 MB payload) and its gradient is complete when the heads' backward is, i.e. before the RPN / FPN /
        backbone backward runs.  The remaining slices are reduced in step(); the payload is still reduced exactly once
        (volume-level data parallelism).  Only for one backward per optimizer step (BATCH_SIZE == 1 per
        rank): with gradient accumulation the early launch would reduce a partial sum."""
        if not self._distributed():
            return self
        self._side = torch.cuda.Stream()
        for n in names:
            if n not in self.slices:
                continue
            a, b = self.slices[n]
            self._early.append((a, b))

            def hook(param, a=a, b=b):
                if not self._distributed():
                    return
                ready = torch.cuda.Event()
                ready.record()
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    work = dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._pending.append((a, b, work))
            self._param_of[n].register_post_accumulate_grad_hook(hook)
        return self

    def allreduce_grads(self):
        """the one collective of the data-parallel path (split into the early slice(s) + the rest when overlapped)"""
        if not self._distributed():
            return
        done = sorted((a, b) for a, b, _ in self._pending)
        pos = 0
        for a, b in done + [(self.numel, self.numel)]:
            if a > pos:
                dist.all_reduce(self.flat_grad[pos:a], op=dist.ReduceOp.SUM, group=self.group)
            pos = max(pos, b)
        for _, _, work in self._pending:
            work.wait()                       # the current stream waits for the side-stream collective
        self._pending = []

    def grad_norm(self):
        self.sumsq.zero_()
        ops.sumsq_into(self.flat_grad, self.sumsq)
        return self.sumsq

    def clip_(self):
        """clip_grad_norm_(params, clip_norm) on the accumulated flat gradient, in place, without a host sync (the
        non-stepping iterations of gradient accumulation, reference model.py:1641)"""
        norm = torch.sqrt(self.grad_norm()).float()
        self.flat_grad.mul_(torch.clamp(self.clip_norm / (norm + 1e-6), max=1.0))

    def step(self):
        self.allreduce_grads()
        self.grad_norm()
        ops.sgd_clip_step(self.flat_param, self.flat_grad, self.flat_mom, self.wd_mask, self.sumsq, self.clip_norm, self.lr,
                          self.momentum, self.weight_decay)


def flatten_grads_cpu(params):
    """Device-agnostic piece of the DP path used by the gloo CPU tests: gradient views into one flat buffer."""
    params = [p for p in params if p.requires_grad]
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is not None:
            flat[off:off + k].copy_(p.grad.reshape(-1))
        p.grad = flat[off:off + k].view(p.shape)
        off += k
    return flat


def shard_volumes(num_volumes, rank, world):
    """Volume-level partition: rank r owns volumes r, r+world, ... (independent units, no data-path collective)."""
    return list(range(rank, num_volumes, world))
