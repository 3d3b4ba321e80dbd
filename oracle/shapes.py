"""TEST INFRASTRUCTURE: the checkpoint ABI (state_dict keys and shapes) of reference MaskRCNN
(model.py:1259-1304, backbone.py, mask_branch.py), written out as a table so that tests can build
deterministic weights without importing the reference.  SURVEY.md 8b: 220 entries at full width."""


def _bn(shapes, p, c):
    shapes[p + ".weight"] = (c,); shapes[p + ".bias"] = (c,)
    shapes[p + ".running_mean"] = (c,); shapes[p + ".running_var"] = (c,)
    shapes[p + ".num_batches_tracked"] = ()


def _conv(shapes, p, co, ci, k, bias=True):
    k = (k, k, k) if isinstance(k, int) else tuple(k)
    shapes[p + ".weight"] = (co, ci) + k
    if bias:
        shapes[p + ".bias"] = (co,)


def backbone_shapes_table(shapes, prefix="fpn", channels=(16, 32), layers=(2, 3), stem_k=(3, 7, 7), in_ch=1):
    _conv(shapes, prefix + ".C1.0", channels[0], in_ch, stem_k)
    _bn(shapes, prefix + ".C1.1", channels[0])
    inplanes = channels[0]
    for si, (planes, nblocks) in enumerate(zip(channels, layers)):
        for b in range(nblocks):
            p = "%s.C%d.%d" % (prefix, si + 2, b)
            expand = b == 0
            _conv(shapes, p + ".conv1", planes, inplanes, 1); _bn(shapes, p + ".bn1", planes)
            _conv(shapes, p + ".conv2", planes, planes, (1, 3, 3)); _bn(shapes, p + ".bn2", planes)
            _conv(shapes, p + ".conv3", planes, planes, (3, 1, 1)); _bn(shapes, p + ".bn3", planes)
            if expand:
                _conv(shapes, p + ".conv4", planes * 4, planes, 1); _bn(shapes, p + ".bn4", planes * 4)
                _conv(shapes, p + ".downsample.0", planes * 4, inplanes, 1); _bn(shapes, p + ".downsample.1", planes * 4)
                inplanes = planes * 4
            else:
                _conv(shapes, p + ".conv4", inplanes, planes, 1); _bn(shapes, p + ".bn4", inplanes)


def unet_shapes_table(shapes, prefix, in_ch, n_classes, b):
    c = lambda n, co, ci, k: _conv(shapes, prefix + "." + n, co, ci, k, bias=False)
    c("conv3d_c1_1", b, in_ch, 3); c("conv3d_c1_2", b, b, 3); c("lrelu_conv_c1.1", b, b, 3)
    for lvl, m in ((2, 2), (3, 4), (4, 8), (5, 16)):
        c("conv3d_c%d" % lvl, b * m, b * m // 2, 3)
        c("norm_lrelu_conv_c%d.2" % lvl, b * m, b * m, 3)
    c("norm_lrelu_upscale_conv_norm_lrelu_l0.3", b * 8, b * 16, 3)
    c("conv3d_l0", b * 8, b * 8, 1)
    c("conv_norm_lrelu_l1.0", b * 16, b * 16, 3); c("conv3d_l1", b * 8, b * 16, 1)
    c("norm_lrelu_upscale_conv_norm_lrelu_l1.3", b * 4, b * 8, 3)
    c("conv_norm_lrelu_l2.0", b * 8, b * 8, 3); c("conv3d_l2", b * 4, b * 8, 1)
    c("norm_lrelu_upscale_conv_norm_lrelu_l2.3", b * 2, b * 4, 3)
    c("conv_norm_lrelu_l3.0", b * 4, b * 4, 3); c("conv3d_l3", b * 2, b * 4, 1)
    c("norm_lrelu_upscale_conv_norm_lrelu_l3.3", b, b * 2, 3)
    c("conv_norm_lrelu_l4.0", b * 2, b * 2, 3); c("conv3d_l4", n_classes, b * 2, 1)
    c("ds2_1x1_conv3d", n_classes, b * 8, 1); c("ds3_1x1_conv3d", n_classes, b * 4, 1)
    c("out_upscale_conv.1", n_classes, n_classes, 5)


def maskrcnn_shapes(fpn=128, rpn=256, unet=20, fc=128, pool=12, num_classes=8, channels=(16, 32)):
    s = {}
    backbone_shapes_table(s, "fpn", channels)
    _conv(s, "fpn.P3_conv1", fpn, channels[1] * 4, 1); _conv(s, "fpn.P3_conv2", fpn, fpn, 3)
    _conv(s, "fpn.P2_conv1", fpn, channels[0] * 4, 1); _conv(s, "fpn.P2_conv2", fpn, fpn, 3)
    _conv(s, "rpn.conv_shared", rpn, fpn, 3); _conv(s, "rpn.conv_class", 2, rpn, 1); _conv(s, "rpn.conv_bbox", 6, rpn, 1)
    _conv(s, "classifier.conv1", fc, fpn, pool); _bn(s, "classifier.bn1", fc)
    _conv(s, "classifier.conv2", fc, fc, 1); _bn(s, "classifier.bn2", fc)
    s["classifier.linear_class.weight"] = (2, fc); s["classifier.linear_class.bias"] = (2,)
    s["classifier.linear_bbox.weight"] = (12, fc); s["classifier.linear_bbox.bias"] = (12,)
    unet_shapes_table(s, "mask.modified_u_net", 1, num_classes, unet)
    return s
