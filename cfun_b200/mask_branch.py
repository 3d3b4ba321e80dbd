"""Modified 3-D U-Net mask branch on the cfun_b200 CUDA ops.

Same class name, constructor signature, attribute names and state_dict keys as reference mask_branch.py:11-122; the
forward (mask_branch.py:124-220) is re-expressed over fused passes: every InstanceNorm3d -> LeakyReLU (-> nearest x2)
chain, including a preceding Dropout3d channel mask, is one statistics kernel + one apply kernel (ops.instnorm_lrelu),
and every Conv3d is ops.conv3d; where a conv feeds a norm directly, the statistics come out of the conv's epilogue."""
import torch
import torch.nn as nn

from . import ops
from .layers import Conv3d, Slot


class Modified3DUNet(nn.Module):
    def __init__(self, in_channels, n_classes, stage, base_n_filter=32):
        super().__init__()
        self.in_channels = in_channels
        self.n_classes = n_classes
        self.base_n_filter = base_n_filter
        self.stage = stage
        b = base_n_filter
        self.lrelu = Slot("LeakyReLU(0.01)")
        self.dropout3d = Slot("Dropout3d(p=0.6)")
        self.upsacle = Slot("Upsample(x2, nearest)")
        self.dropout_p = 0.6
        self.use_dropout = True          # LiTS variant has none (LiTS_2017/mask_branch.py:19,130)
        self.injected_drop = None        # tests: list of 5 per-call channel masks [N,C,1,1,1] (already scaled)

        c3 = lambda ci, co, s=1: Conv3d(ci, co, kernel_size=3, stride=s, padding=1, bias=False)
        c1 = lambda ci, co: Conv3d(ci, co, kernel_size=1, stride=1, padding=0, bias=False)
        self.conv3d_c1_1 = c3(in_channels, b)
        self.conv3d_c1_2 = c3(b, b)
        self.lrelu_conv_c1 = nn.Sequential(Slot("LeakyReLU"), c3(b, b))
        self.inorm3d_c1 = Slot("InstanceNorm3d")
        for lvl, m in ((2, 2), (3, 4), (4, 8), (5, 16)):
            setattr(self, "conv3d_c%d" % lvl, c3(b * m // 2, b * m, 2))
            setattr(self, "norm_lrelu_conv_c%d" % lvl, nn.Sequential(Slot("InstanceNorm3d"), Slot("LeakyReLU"), c3(b * m, b * m)))
            if lvl < 5:
                setattr(self, "inorm3d_c%d" % lvl, Slot("InstanceNorm3d"))
        up_block = lambda ci, co: nn.Sequential(Slot("InstanceNorm3d"), Slot("LeakyReLU"), Slot("Upsample x2"), c3(ci, co),
                                                Slot("InstanceNorm3d"), Slot("LeakyReLU"))
        cnl = lambda ci, co: nn.Sequential(c3(ci, co), Slot("InstanceNorm3d"), Slot("LeakyReLU"))
        self.norm_lrelu_upscale_conv_norm_lrelu_l0 = up_block(b * 16, b * 8)
        self.conv3d_l0 = c1(b * 8, b * 8)
        self.inorm3d_l0 = Slot("InstanceNorm3d")
        self.conv_norm_lrelu_l1 = cnl(b * 16, b * 16)
        self.conv3d_l1 = c1(b * 16, b * 8)
        self.norm_lrelu_upscale_conv_norm_lrelu_l1 = up_block(b * 8, b * 4)
        self.conv_norm_lrelu_l2 = cnl(b * 8, b * 8)
        self.conv3d_l2 = c1(b * 8, b * 4)
        self.norm_lrelu_upscale_conv_norm_lrelu_l2 = up_block(b * 4, b * 2)
        self.conv_norm_lrelu_l3 = cnl(b * 4, b * 4)
        self.conv3d_l3 = c1(b * 4, b * 2)
        self.norm_lrelu_upscale_conv_norm_lrelu_l3 = up_block(b * 2, b)
        self.conv_norm_lrelu_l4 = cnl(b * 2, b * 2)
        self.conv3d_l4 = c1(b * 2, n_classes)
        self.ds2_1x1_conv3d = c1(b * 8, n_classes)
        self.ds3_1x1_conv3d = c1(b * 4, n_classes)
        self.out_upscale_conv = nn.Sequential(Slot("Upsample x2"),
                                              Conv3d(n_classes, n_classes, kernel_size=5, stride=1, padding=2, bias=False))

    # -- Dropout3d draws: per (sample, channel) keep mask scaled by 1/(1-p) (mask_branch.py:19) ---------------
    def _drop_masks(self, n, device):
        if not (self.training and self.use_dropout):
            return [None] * 5
        if self.injected_drop is not None:
            return [m[:n].reshape(min(n, m.shape[0]), -1).to(device=device, dtype=torch.float32) for m in self.injected_drop]
        b, keep = self.base_n_filter, 1.0 - self.dropout_p
        return [torch.bernoulli(torch.full((n, b * m), keep, device=device)) / keep for m in (1, 2, 4, 8, 16)]

    def forward(self, x):
        IN = ops.instnorm_lrelu

        def CIN(conv, t, drop=None, t2=None):     # IN(conv(cat(t, t2)), drop): one autograd node where the conv has the fused backward
            return ops.conv_in_lrelu(t, conv.weight, conv.bias, conv.stride, conv.padding, drop, x2=t2)
        drops = self._drop_masks(x.shape[0], x.device)
        # level 1 context (mask_branch.py:125-136)
        out = self.conv3d_c1_1(x)
        residual_1 = out
        # lrelu -> conv and dropout -> lrelu -> conv: the activation (and the Dropout3d channel scale) are applied on the way
        # into the conv's operand pack (ops.lrelu_conv3d)
        c12, c13 = self.conv3d_c1_2, self.lrelu_conv_c1[1]
        out = ops.lrelu_conv3d(out, c12.weight, c12.bias, c12.stride, c12.padding)
        out = ops.lrelu_conv3d(out, c13.weight, c13.bias, c13.stride, c13.padding, scale=drops[0])
        context_1, out = ops.add_lrelu_instnorm(out, residual_1)     # s = out + residual_1 -> (lrelu(s), IN(s)) as one node
        # levels 2..5 context (mask_branch.py:138-183)
        ctx = {}
        for lvl in (2, 3, 4, 5):
            out = getattr(self, "conv3d_c%d" % lvl)(out, in_stats=True)
            residual = out
            conv = getattr(self, "norm_lrelu_conv_c%d" % lvl)[2]
            out = CIN(conv, IN(out), drop=drops[lvl - 1])     # conv -> dropout -> norm -> lrelu as one fused node
            out = conv(out)
            out = out + residual
            if lvl < 5:
                out = IN(out)
                ctx[lvl] = out

        def up_block(seq, t):   # norm -> lrelu -> upsample x2 -> conv -> norm -> lrelu
            conv = seq[3]
            if conv.stride == (1, 1, 1) and conv.padding == (1, 1, 1):
                return ops.upnorm_conv_in_lrelu(t, conv.weight, conv.bias)     # the upsampled tensor is never materialised
            return CIN(conv, IN(t, up=2))

        out = up_block(self.norm_lrelu_upscale_conv_norm_lrelu_l0, out)
        out = IN(self.conv3d_l0(out, in_stats=True))
        out = CIN(self.conv_norm_lrelu_l1[0], out, t2=ctx[4])        # conv(cat(out, skip)): the pack is built from the two tensors
        out = self.conv3d_l1(out)
        out = up_block(self.norm_lrelu_upscale_conv_norm_lrelu_l1, out)
        out = CIN(self.conv_norm_lrelu_l2[0], out, t2=ctx[3])        # conv(cat(out, skip)): the pack is built from the two tensors
        ds2 = out
        out = self.conv3d_l2(out)
        out = up_block(self.norm_lrelu_upscale_conv_norm_lrelu_l2, out)
        out = CIN(self.conv_norm_lrelu_l3[0], out, t2=ctx[2])        # conv(cat(out, skip)): the pack is built from the two tensors
        ds3 = out
        out = self.conv3d_l3(out)
        out = up_block(self.norm_lrelu_upscale_conv_norm_lrelu_l3, out)
        out = CIN(self.conv_norm_lrelu_l4[0], out, t2=context_1)        # conv(cat(out, skip)): the pack is built from the two tensors
        out_pred = self.conv3d_l4(out)
        # deep supervision (mask_branch.py:209-215)
        s = ops.upsample2x(self.ds2_1x1_conv3d(ds2)) + self.ds3_1x1_conv3d(ds3)
        out = out_pred + ops.upsample2x(s)
        if self.stage == 'finetune':   # mask_branch.py:216-218
            up = ops.upsample2x(out)
            out = up + self.out_upscale_conv[1](up)
        return out
