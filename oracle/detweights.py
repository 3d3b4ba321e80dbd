"""TEST INFRASTRUCTURE: deterministic, formula-defined weights keyed by state_dict name.

Goldens would otherwise have to carry the 41 M-parameter state_dict.  Both the golden generator (which
loads them into the unmodified reference modules) and the tests (which load them into the oracle and the
CUDA modules) call det_state() with the same shapes and seed, so only the seed is stored.
"""
import math
import zlib
import torch


def det_tensor(key, shape, seed):
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    if key.endswith("running_var"):
        return torch.rand(shape, generator=g) + 0.5
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    is_bn = ".bn" in key or ".C1.1." in key or "downsample.1" in key
    if key.endswith(".bias"):
        return torch.randn(shape, generator=g) * (0.1 if is_bn else 0.05)
    if is_bn:
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 1.0
    if "rpn.conv_bbox" in key or "rpn.conv_class" in key or ".linear_" in key:
        gain = 0.02      # keep RPN deltas / scores small so proposals stay near their anchors
    return torch.randn(shape, generator=g) * (gain * math.sqrt(1.0 / max(fan_in, 1)))


def det_state(shapes, seed):
    """shapes: {key: shape}.  Returns {key: tensor}."""
    return {k: det_tensor(k, s, seed) for k, s in shapes.items()}
