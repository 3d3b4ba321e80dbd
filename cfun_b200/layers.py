"""Parameter-holding layer classes that keep the reference's state_dict keys / shapes (the checkpoint ABI,
reference model.py:1329-1339) while routing compute to the CUDA ops in cfun_b200.ops."""
import torch
import torch.nn as nn

from . import ops


class Conv3d(nn.Conv3d):
    """nn.Conv3d parameter layout (Cout, Cin, kD, kH, kW); forward = hand-written sm_100a conv (ops.conv3d)."""

    def forward(self, x, relu=False, in_stats=False):
        """in_stats: the caller feeds the result straight to ops.instnorm_lrelu (see ops.conv3d)"""
        if self.dilation != (1, 1, 1) or self.groups != 1 or self.padding_mode != "zeros":
            raise RuntimeError("cfun_b200.Conv3d supports dilation=1, groups=1, zero padding (all the reference uses)")
        return ops.conv3d(x, self.weight, self.bias, self.stride, self.padding, relu, in_stats)


class FrozenBatchNorm3d(nn.BatchNorm3d):
    """BatchNorm3d as the reference runs it on this path: always eval, parameters frozen (model.py:1297-1304,
    1401-1406).  Exposes the folded per-channel scale / shift; the affine + ReLU (+ residual) pass is ops.affine_act."""

    def coeffs(self):
        key = (self.weight._version, self.bias._version, self.running_mean._version, self.running_var._version,
               self.weight.device)
        if getattr(self, "_coef_key", None) != key:
            with torch.no_grad():
                a = self.weight / torch.sqrt(self.running_var + self.eps)
                b = self.bias - self.running_mean * a
            self._coef = (a.contiguous(), b.contiguous())
            self._coef_key = key
        return self._coef

    def forward(self, x, relu=False, residual=None):
        a, b = self.coeffs()
        return ops.affine_act(x, a, b, residual, 0.0 if relu else 1.0, 1)


class Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential child indices identical to the reference (so that keys such as
    'norm_lrelu_conv_c2.2.weight' line up); the owning module fuses what the slot stands for."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what

    def forward(self, x):
        raise RuntimeError("Slot('%s') is fused by its parent module and is never called on its own" % self.what)
