"""torch.autograd Functions over the C ABI of libcfun_b200.so.

Host code stays PyTorch (device memory, streams, autograd bookkeeping); every op body is a hand-written sm_100a CUDA
kernel reached through ctypes with raw device pointers and the current CUDA stream.  There is no CPU path: tensors must
live on a CUDA device, and a missing / failing library raises.

Activation layout: logical [N,C,D,H,W] tensors whose memory is N,D,H,W,C (torch.channels_last_3d).
"""
import ctypes as C
import os

import torch
from torch.autograd import Function

from ._lib import lib, check, ConvDesc, f6

ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_TC1 = 0, 1, 2, 3
PASS_FWD, PASS_BWD_DATA, PASS_BWD_WEIGHT = 0, 1, 2
EPI_BIAS, EPI_RELU = 1, 2

# launch accounting for bench.py ("gpu_launches": kernels of OUR library launched in the timed region)
_calls = {"n": 0}


def call_count():
    return _calls["n"]


def tc_debug_status():
    """(site, blockIdx.x, blockIdx.y, threadIdx.x, parity, spins) of the first tcgen05 pipeline wait that timed out, or None"""
    buf = (C.c_int * 8)()
    check(lib.cfun_tc_debug_status(buf), "cfun_tc_debug_status")
    vals = list(buf)
    return None if vals[0] == 0 else tuple(vals[:6])


def kernel_timing(on):
    """bracket the main kernel of every tensor-core conv call with CUDA events (bench.py rooflines)"""
    check(lib.cfun_kernel_timing(1 if on else 0), "cfun_kernel_timing")


def last_kernel_ms():
    ms = C.c_float(0)
    check(lib.cfun_last_kernel_ms(C.byref(ms)), "cfun_last_kernel_ms")
    return float(ms.value)


def launch_count():
    """CUDA kernels launched by libcfun_b200.so so far in this process"""
    return int(lib.cfun_launch_count())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cfun_b200 ops are CUDA-only (sm_100a); got a %s tensor. There is no CPU fallback."
                               % t.device)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def is_cl(x):
    """memory order is exactly N,D,H,W,C and dense"""
    return x.permute(0, 2, 3, 4, 1).is_contiguous()


def to_cl(x):
    if x.dtype != torch.float32:
        x = x.float()
    if is_cl(x):
        return x
    return x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


def empty_cl(N, Cc, D, H, W, device, dtype=torch.float32):
    return torch.empty((N, D, H, W, Cc), device=device, dtype=dtype).permute(0, 4, 1, 2, 3)


def zeros_cl(N, Cc, D, H, W, device, dtype=torch.float32):
    return torch.zeros((N, D, H, W, Cc), device=device, dtype=dtype).permute(0, 4, 1, 2, 3)


_ws = {}
_ws_gen = {"n": 0}


def workspace(nbytes, device):
    """Stream-ordered scratch reused by consecutive calls on the same device (grown geometrically).  Every reallocation
    bumps workspace_generation(): anything that baked the old address in (captured CUDA graphs) must be rebuilt."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25) + 4096, 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
        _ws_gen["n"] += 1
    return buf


def workspace_generation():
    return _ws_gen["n"]


_prof = {"on": False, "rows": []}


def profile_start():
    """Per-call device timing of every library call (CUDA events on the current stream); for tools/, never for bench."""
    _prof["on"], _prof["rows"] = True, []


def profile_stop():
    """-> list of (entry point, tag, milliseconds); tag carries the conv geometry and the algorithm picked."""
    _prof["on"] = False
    torch.cuda.synchronize()
    return [(n, t, a.elapsed_time(b)) for n, t, a, b in _prof["rows"]]


def _conv_tag(d, pass_id, algo):
    picked = lib.cfun_conv3d_pick_algo(C.byref(d), pass_id) if algo == ALGO_AUTO else algo
    return "N%d %d->%d in%dx%dx%d k%d%d%d s%d%d%d algo%d" % (d.N, d.Cin, d.Cout, d.Din, d.Hin, d.Win, d.kD, d.kH, d.kW,
                                                          d.sD, d.sH, d.sW, picked)


_CAPTURE_DEBUG = bool(int(os.environ.get("CFUN_DEBUG_CAPTURE", "0")))


def _capture_state():
    """0 = not capturing, 1 = capturing, 2 = capture invalidated (bring-up aid, CFUN_DEBUG_CAPTURE=1)"""
    return int(lib.cfun_stream_capture_status(_stream()))


def _run(name, *args, tag=""):
    _calls["n"] += 1
    if _CAPTURE_DEBUG:
        before = _capture_state()
        rc = getattr(lib, name)(*args)
        after = _capture_state()
        if before == 2 or after == 2:
            raise RuntimeError("stream capture invalidated %s %s (rc=%d, %s)" % ("before" if before == 2 else "inside", name, rc, tag))
        check(rc, name)
        return
    if _prof["on"]:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(getattr(lib, name)(*args), name)
        b.record()
        _prof["rows"].append((name, tag, a, b))
        return
    check(getattr(lib, name)(*args), name)


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(int(t) for t in v)


# --------------------------------------------------------------------------------------------------------
# conv3d
# --------------------------------------------------------------------------------------------------------
def _conv_desc(x_shape, w_shape, stride, padding):
    N, Cin, D, H, W = x_shape
    Cout, Cin_w, kD, kH, kW = w_shape
    if Cin_w != Cin:
        raise RuntimeError("conv3d: weight expects %d input channels, got %d" % (Cin_w, Cin))
    s, p = _triple(stride), _triple(padding)
    Do = (D + 2 * p[0] - kD) // s[0] + 1
    Ho = (H + 2 * p[1] - kH) // s[1] + 1
    Wo = (W + 2 * p[2] - kW) // s[2] + 1
    if min(Do, Ho, Wo) <= 0:
        raise RuntimeError("conv3d: kernel %s larger than padded input %s" % ((kD, kH, kW), (D, H, W)))
    return ConvDesc(N, Cin, D, H, W, Cout, Do, Ho, Wo, kD, kH, kW, s[0], s[1], s[2], p[0], p[1], p[2])


def conv3d_supported(x_shape, w_shape, stride, padding, pass_id, algo):
    d = _conv_desc(tuple(x_shape), tuple(w_shape), stride, padding)
    return bool(lib.cfun_conv3d_supported(C.byref(d), pass_id, algo))


_default_algo = {"algo": ALGO_AUTO}


def set_conv_algo(algo):
    """ALGO_AUTO (default): tcgen05 where supported else CUDA cores; ALGO_SIMT / ALGO_TC / ALGO_TC1 force one."""
    _default_algo["algo"] = algo


class Conv3dFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding, relu, stats_out=None):
        _require_cuda(x, w, b)
        x = to_cl(x)
        w = w.contiguous()
        d = _conv_desc(x.shape, w.shape, stride, padding)
        y = empty_cl(d.N, d.Cout, d.Dout, d.Hout, d.Wout, x.device)
        algo = _default_algo["algo"]
        ws_bytes = lib.cfun_conv3d_workspace_size(C.byref(d), PASS_FWD, algo)
        ws = workspace(ws_bytes, x.device)
        epi = (EPI_BIAS if b is not None else 0) | (EPI_RELU if relu else 0)
        # Fused backward (conv_fused.cu): where all three passes run on the halo-family tcgen05 kernels, the forward keeps its
        # split-bf16 pack of x for the weight gradient, and the backward packs dy once for both gradients.
        pack_bytes = lib.cfun_conv3d_pack_bytes(C.byref(d)) if (algo == ALGO_AUTO and ctx.needs_input_grad[1]) else 0
        xpack = None
        if pack_bytes and stats_out is not None:
            # the consumer is an InstanceNorm: its per-(sample, channel) sums come out of the conv epilogue (stats_out is a
            # one-element list the caller reads the accumulator from; it stays empty on the other dispatch paths)
            xpack = torch.empty(pack_bytes, dtype=torch.uint8, device=x.device)
            acc = torch.empty(2 * d.N * d.Cout, dtype=torch.float64, device=x.device)
            _run("cfun_conv3d_fwd_stats", C.byref(d), _ptr(x), _ptr(w), _ptr(b), _ptr(y), epi, _ptr(xpack), xpack.numel(),
                 _ptr(acc), _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_FWD, algo) if _prof["on"] else "")
            stats_out.append(acc)
        elif pack_bytes:
            xpack = torch.empty(pack_bytes, dtype=torch.uint8, device=x.device)
            _run("cfun_conv3d_fwd_keep_pack", C.byref(d), _ptr(x), _ptr(w), _ptr(b), _ptr(y), epi, _ptr(xpack), xpack.numel(),
                 _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_FWD, algo) if _prof["on"] else "")
        else:
            _run("cfun_conv3d_fwd", C.byref(d), _ptr(x), _ptr(w), _ptr(b), _ptr(y), epi, algo, _ptr(ws), ws.numel(), _stream(),
                 tag=_conv_tag(d, PASS_FWD, algo) if _prof["on"] else "")
        ctx.d = d
        ctx.relu = relu
        ctx.has_bias = b is not None
        ctx.algo = algo
        ctx.fused = xpack is not None
        # the fused backward reads x only through its pack
        ctx.save_for_backward(xpack if xpack is not None else x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        d = ctx.d
        dy = to_cl(dy)
        if ctx.relu:
            dy = to_cl(torch.where(y > 0, dy, torch.zeros((), device=dy.device)))
        dx = dw = db = None
        if ctx.fused:
            xpack = x
            if ctx.needs_input_grad[0]:
                dx = empty_cl(d.N, d.Cin, d.Din, d.Hin, d.Win, dy.device)
            if ctx.needs_input_grad[1]:
                dw = torch.empty_like(w)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = torch.empty(d.Cout, device=dy.device)
            ws_bytes = lib.cfun_conv3d_bwd_fused_workspace_size(C.byref(d))
            ws = workspace(ws_bytes, dy.device)
            _run("cfun_conv3d_bwd_fused", C.byref(d), _ptr(xpack), xpack.numel(), _ptr(dy), _ptr(w), _ptr(dx), _ptr(dw), _ptr(db),
                 _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_BWD_WEIGHT, ctx.algo) if _prof["on"] else "")
            return dx, dw, db, None, None, None, None
        if ctx.needs_input_grad[0]:
            dx = empty_cl(d.N, d.Cin, d.Din, d.Hin, d.Win, dy.device)
            ws_bytes = lib.cfun_conv3d_workspace_size(C.byref(d), PASS_BWD_DATA, ctx.algo)
            ws = workspace(ws_bytes, dy.device)
            _run("cfun_conv3d_bwd_data", C.byref(d), _ptr(dy), _ptr(w), _ptr(dx), ctx.algo, _ptr(ws), ws.numel(), _stream(),
                 tag=_conv_tag(d, PASS_BWD_DATA, ctx.algo) if _prof["on"] else "")
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty_like(w)
            db = torch.empty(d.Cout, device=dy.device) if ctx.has_bias else None
            ws_bytes = lib.cfun_conv3d_workspace_size(C.byref(d), PASS_BWD_WEIGHT, ctx.algo)
            ws = workspace(ws_bytes, dy.device)
            _run("cfun_conv3d_bwd_weight", C.byref(d), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), ctx.algo, _ptr(ws),
                 ws.numel(), _stream(), tag=_conv_tag(d, PASS_BWD_WEIGHT, ctx.algo) if _prof["on"] else "")
        return dx, dw, db, None, None, None, None


def conv3d(x, w, b=None, stride=1, padding=0, relu=False, in_stats=False):
    """in_stats=True: the output feeds an InstanceNorm (ops.instnorm_lrelu); where the conv runs on the halo-family tcgen05
    kernels its epilogue accumulates the norm's statistics, which ride along on the returned tensor (y._cfun_in_stats)
    and save the norm its own pass over y."""
    if not in_stats:
        return Conv3dFn.apply(x, w, b, stride, padding, relu)
    holder = []
    y = Conv3dFn.apply(x, w, b, stride, padding, relu, holder)
    if holder:
        y._cfun_in_stats = holder[0]
    return y


class PreActConv3dFn(Function):
    """conv3d(leaky_relu(x * scale[n,c], slope), w) with the activation applied on the way into the conv's operand pack (the
    activated tensor is never written): the U-Net's conv3d_c1_2(lrelu(out)) and lrelu_conv_c1(lrelu(dropout3d(out)))
    (reference mask_branch.py:127-131).  Backward = the fused conv backward on the kept pack, then the activation backward."""

    @staticmethod
    def forward(ctx, x, w, scale, slope, stride, padding):
        _require_cuda(x, w, scale)
        x = to_cl(x)
        w = w.contiguous()
        d = _conv_desc(x.shape, w.shape, stride, padding)
        if scale is not None:
            scale = scale.reshape(d.N, d.Cin).float().contiguous()
        y = empty_cl(d.N, d.Cout, d.Dout, d.Hout, d.Wout, x.device)
        ws = workspace(lib.cfun_conv3d_workspace_size(C.byref(d), PASS_FWD, ALGO_AUTO), x.device)
        xpack = torch.empty(lib.cfun_conv3d_pack_bytes(C.byref(d)), dtype=torch.uint8, device=x.device)
        _run("cfun_conv3d_fwd_keep_pack_preact", C.byref(d), _ptr(x), _ptr(scale), float(slope), _ptr(w), None, _ptr(y), 0, _ptr(xpack),
             xpack.numel(), _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_FWD, ALGO_AUTO) if _prof["on"] else "")
        ctx.save_for_backward(xpack, w, x, scale)
        ctx.d, ctx.slope = d, float(slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        xpack, w, x, scale = ctx.saved_tensors
        d = ctx.d
        dy = to_cl(dy)
        dev = dy.device
        da = empty_cl(d.N, d.Cin, d.Din, d.Hin, d.Win, dev) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        ws = workspace(lib.cfun_conv3d_bwd_fused_workspace_size(C.byref(d)), dev)
        _run("cfun_conv3d_bwd_fused", C.byref(d), _ptr(xpack), xpack.numel(), _ptr(dy), _ptr(w), _ptr(da), _ptr(dw), None,
             _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_BWD_WEIGHT, ALGO_AUTO) if _prof["on"] else "")
        dx = None
        if da is not None:      # gradient of the activated input -> gradient of x
            dx = empty_cl(d.N, d.Cin, d.Din, d.Hin, d.Win, dev)
            zero = torch.zeros_like(scale) if scale is not None else None
            _run("cfun_affine_act_bwd", _ptr(x), _ptr(scale), _ptr(zero), d.Cin if scale is not None else 0, None, _ptr(da), _ptr(dx),
                 None, None, d.N, d.Din, d.Hin, d.Win, d.Cin, d.Cin, 0, 1, ctx.slope, _stream())
        return dx, dw, None, None, None, None


def lrelu_conv3d(x, w, b=None, stride=1, padding=0, scale=None, slope=0.01):
    """conv3d(leaky_relu(x * scale, slope), w, b): fused into the conv's operand pack where the conv has the fused tcgen05
    backward and no bias (PreActConv3dFn), else the two separate ops.  scale: None or a per-(sample, channel) factor [N,C]."""
    if b is None and _default_algo["algo"] == ALGO_AUTO and x.is_cuda and w.requires_grad and torch.is_grad_enabled():
        d = _conv_desc(tuple(x.shape), tuple(w.shape), stride, padding)
        if lib.cfun_conv3d_pack_bytes(C.byref(d)) and lib.cfun_conv3d_preact_supported(C.byref(d)):
            return PreActConv3dFn.apply(x, w, scale, slope, stride, padding)
    a = leaky_relu(x, slope) if scale is None else affine_act(x, scale, torch.zeros_like(scale), None, slope, 1)
    return conv3d(a, w, b, stride, padding)


class FcConvFn(Function):
    """Conv3d whose kernel covers the whole (un-padded) input: Classifier.conv1 (model.py:758).  x is consumed in
    NCDHW-contiguous order so K = (ci, kd, kh, kw) matches the checkpoint weight layout with no repack."""

    @staticmethod
    def forward(ctx, x, w, b):
        _require_cuda(x, w, b)
        x = x.contiguous()
        w = w.contiguous()
        M, K, Nout = x.shape[0], x[0].numel() if x.shape[0] else w[0].numel(), w.shape[0]
        y = torch.empty((M, Nout), device=x.device)
        _run("cfun_fc_fwd", M, Nout, K, _ptr(x), _ptr(w), _ptr(b), _ptr(y), _stream())
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y.view(M, Nout, 1, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        M, Nout = x.shape[0], w.shape[0]
        K = w[0].numel()
        dy = dy.reshape(M, Nout).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _run("cfun_fc_bwd_data", M, Nout, K, _ptr(dy), _ptr(w), _ptr(dx), _stream())
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            db = torch.empty(Nout, device=dy.device) if ctx.has_bias else None
            _run("cfun_fc_bwd_weight", M, Nout, K, _ptr(dy), _ptr(x), _ptr(dw), _ptr(db), _stream())
        return dx, dw, db


def fc_conv(x, w, b=None):
    return FcConvFn.apply(x, w, b)


# --------------------------------------------------------------------------------------------------------
# affine / activation / norm
# --------------------------------------------------------------------------------------------------------
class AffineActFn(Function):
    """y = leaky_relu(x * a + b (+ r), slope), optionally nearest-upsampled x2.  a, b: [C] or [N,C] constants."""

    @staticmethod
    def forward(ctx, x, a, b, r, slope, up):
        _require_cuda(x, a, b, r)
        x = to_cl(x)
        N, Cc, D, H, W = x.shape
        if r is not None:
            r = to_cl(r)
        nstride = 0
        if a is not None:
            a = a.contiguous().float()
            b = b.contiguous().float()
            nstride = Cc if a.dim() == 2 else 0
        y = empty_cl(N, Cc, D * up, H * up, W * up, x.device)
        _run("cfun_affine_act_fwd", _ptr(x), _ptr(a), _ptr(b), nstride, _ptr(r), _ptr(y), N, D, H, W, Cc, Cc, 0, up,
             float(slope), _stream())
        ctx.save_for_backward(x, a, b, r)
        ctx.cfg = (nstride, float(slope), up)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a, b, r = ctx.saved_tensors
        nstride, slope, up = ctx.cfg
        N, Cc, D, H, W = x.shape
        dy = to_cl(dy)
        dx = empty_cl(N, Cc, D, H, W, x.device)
        dr = empty_cl(N, Cc, D, H, W, x.device) if (r is not None and ctx.needs_input_grad[3]) else None
        _run("cfun_affine_act_bwd", _ptr(x), _ptr(a), _ptr(b), nstride, _ptr(r), _ptr(dy), _ptr(dx), _ptr(dr), None, N, D,
             H, W, Cc, Cc, 0, up, slope, _stream())
        return dx, None, None, dr, None, None


def affine_act(x, a=None, b=None, r=None, slope=0.0, up=1):
    return AffineActFn.apply(x, a, b, r, slope, up)


def leaky_relu(x, slope=0.01):
    return AffineActFn.apply(x, None, None, None, slope, 1)


class Cat2Fn(Function):
    """torch.cat([a, b], dim=1) for channels-last tensors as one vectorised pass (and one split pass backward)."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        a, b = to_cl(a), to_cl(b)
        N, C1, D, H, W = a.shape
        C2 = b.shape[1]
        out = empty_cl(N, C1 + C2, D, H, W, a.device)
        _run("cfun_cat2_channels", _ptr(a), C1, _ptr(b), C2, _ptr(out), N * D * H * W, _stream())
        ctx.dims = (N, C1, C2, D, H, W)
        return out

    @staticmethod
    def backward(ctx, dy):
        N, C1, C2, D, H, W = ctx.dims
        dy = to_cl(dy)
        da = empty_cl(N, C1, D, H, W, dy.device)
        db = empty_cl(N, C2, D, H, W, dy.device)
        _run("cfun_split2_channels", _ptr(dy), C1, C2, _ptr(da), _ptr(db), N * D * H * W, _stream())
        return da, db


def cat_channels(a, b):
    if a.shape[1] % 4 or b.shape[1] % 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        return torch.cat([a, b], dim=1)
    return Cat2Fn.apply(a, b)


def upsample2x(x):
    return AffineActFn.apply(x, None, None, None, 1.0, 2)


class InstNormActFn(Function):
    """InstanceNorm3d(affine=False, eps) of (x * drop) followed by LeakyReLU(slope) and optional nearest x2 upsampling,
    as one statistics pass + one apply pass.  drop: None or per-(n,c) Dropout3d scale [N,C] (0 or 1/(1-p))."""

    @staticmethod
    def forward(ctx, x, drop, eps, slope, up, acc=None):
        _require_cuda(x, drop)
        x = to_cl(x)
        N, Cc, D, H, W = x.shape
        S = D * H * W
        mean = torch.empty((N, Cc), device=x.device)
        rstd = torch.empty((N, Cc), device=x.device)
        if acc is not None:      # sums accumulated by the producing conv's epilogue (ops.conv3d(in_stats=True))
            assert acc.dtype == torch.float64 and acc.numel() == 2 * N * Cc
            _run("cfun_instnorm_finalize", _ptr(acc), N, S, Cc, float(eps), _ptr(mean), _ptr(rstd), _stream())
        else:
            acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=x.device)
            _run("cfun_instnorm_stats", _ptr(x), N, S, Cc, float(eps), _ptr(acc), _ptr(mean), _ptr(rstd), _stream())
        a, b = _in_coeffs(mean, rstd, drop, N, Cc, eps)
        y = empty_cl(N, Cc, D * up, H * up, W * up, x.device)
        _run("cfun_affine_act_fwd", _ptr(x), _ptr(a), _ptr(b), Cc, None, _ptr(y), N, D, H, W, Cc, Cc, 0, up, float(slope),
             _stream())
        ctx.save_for_backward(x, a, b)
        ctx.cfg = (float(slope), up)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a, b = ctx.saved_tensors
        slope, up = ctx.cfg
        N, Cc, D, H, W = x.shape
        dy = to_cl(dy)
        dx = empty_cl(N, Cc, D, H, W, x.device)
        acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=x.device)
        _run("cfun_affine_act_bwd", _ptr(x), _ptr(a), _ptr(b), Cc, None, _ptr(dy), _ptr(dx), None, _ptr(acc), N, D, H, W,
             Cc, Cc, 0, up, slope, _stream())
        _run("cfun_instnorm_bwd_apply", _ptr(x), _ptr(a), _ptr(b), _ptr(acc), _ptr(dx), N, D * H * W, Cc, _stream())
        return dx, None, None, None, None, None


def _in_coeffs(mean, rstd, drop, N, Cc, eps):
    """per-(sample, channel) scale / shift of InstanceNorm applied to x * drop (drop: None or the Dropout3d channel scale)"""
    if drop is None:
        return rstd.contiguous(), (-mean * rstd).contiguous()
    m = drop.reshape(N, Cc).float()
    var = 1.0 / (rstd * rstd) - eps
    rstd2 = torch.rsqrt(m * m * var + eps)
    return (m * rstd2).contiguous(), (-(m * mean) * rstd2).contiguous()


class ConvInstNormActFn(Function):
    """Conv3d (3^3, stride 1, no bias, on the halo-family tcgen05 kernels) -> [Dropout3d channel scale] -> InstanceNorm3d ->
    LeakyReLU (-> nearest x2) as ONE autograd node, so that the conv output's gradient never exists in fp32:
    forward  = cfun_conv3d_fwd_stats (norm statistics from the conv epilogue) + finalize + one apply pass;
    backward = activation/norm backward statistics pass, then cfun_instnorm_bwd_apply_pack writes the norm's input gradient
    straight into the split-bf16 operand pack that cfun_conv3d_bwd_fused_packed feeds to the data- and weight-gradient
    kernels (the unfused path writes it in fp32 and re-reads it to pack it)."""

    @staticmethod
    def forward(ctx, x, w, drop, stride, padding, eps, slope, up, x2=None):
        """x2: the conv input is torch.cat((x, x2), dim=1) -- packed straight from the two tensors, never materialised"""
        _require_cuda(x, w, drop, x2)
        x = to_cl(x)
        w = w.contiguous()
        c1 = x.shape[1]
        c2 = 0 if x2 is None else x2.shape[1]
        d = _conv_desc((x.shape[0], c1 + c2) + tuple(x.shape[2:]), w.shape, stride, padding)
        N, Cc, D, H, W = d.N, d.Cout, d.Dout, d.Hout, d.Wout
        y = empty_cl(N, Cc, D, H, W, x.device)
        ws = workspace(lib.cfun_conv3d_workspace_size(C.byref(d), PASS_FWD, ALGO_AUTO), x.device)
        xpack = torch.empty(lib.cfun_conv3d_pack_bytes(C.byref(d)), dtype=torch.uint8, device=x.device)
        acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=x.device)
        if x2 is None:
            _run("cfun_conv3d_fwd_stats", C.byref(d), _ptr(x), _ptr(w), None, _ptr(y), 0, _ptr(xpack), xpack.numel(), _ptr(acc),
                 _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_FWD, ALGO_AUTO) if _prof["on"] else "")
        else:
            x2 = to_cl(x2)
            _run("cfun_conv3d_fwd_stats_cat", C.byref(d), _ptr(x), c1, _ptr(x2), c2, _ptr(w), _ptr(y), _ptr(xpack), xpack.numel(),
                 _ptr(acc), _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_FWD, ALGO_AUTO) if _prof["on"] else "")
        ctx.split = (c1, c2)
        mean = torch.empty((N, Cc), device=x.device)
        rstd = torch.empty((N, Cc), device=x.device)
        _run("cfun_instnorm_finalize", _ptr(acc), N, D * H * W, Cc, float(eps), _ptr(mean), _ptr(rstd), _stream())
        a, b = _in_coeffs(mean, rstd, drop, N, Cc, eps)
        z = empty_cl(N, Cc, D * up, H * up, W * up, x.device)
        _run("cfun_affine_act_fwd", _ptr(y), _ptr(a), _ptr(b), Cc, None, _ptr(z), N, D, H, W, Cc, Cc, 0, up, float(slope), _stream())
        ctx.save_for_backward(xpack, w, y, a, b)
        ctx.d, ctx.cfg = d, (float(slope), up)
        return z

    @staticmethod
    def backward(ctx, dz):
        xpack, w, y, a, b = ctx.saved_tensors
        d = ctx.d
        slope, up = ctx.cfg
        N, Cc, D, H, W = d.N, d.Cout, d.Dout, d.Hout, d.Wout
        dz = to_cl(dz)
        dev = dz.device
        g = empty_cl(N, Cc, D, H, W, dev)
        acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=dev)
        _run("cfun_affine_act_bwd", _ptr(y), _ptr(a), _ptr(b), Cc, None, _ptr(dz), _ptr(g), None, _ptr(acc), N, D, H, W, Cc, Cc, 0, up,
             slope, _stream())
        G, P = C.c_int(0), C.c_int(0)
        ybytes = lib.cfun_conv3d_dy_pack_geometry(C.byref(d), C.byref(G), C.byref(P))
        ypack = torch.empty(ybytes, dtype=torch.uint8, device=dev)
        _run("cfun_instnorm_bwd_apply_pack", _ptr(y), _ptr(a), _ptr(b), _ptr(acc), _ptr(g), N, D, H, W, Cc, _ptr(ypack),
             C.c_void_p(ypack.data_ptr() + ybytes // 2), G.value, P.value, _stream())
        del g
        c1, c2 = ctx.split
        want_dx = ctx.needs_input_grad[0] or (c2 > 0 and ctx.needs_input_grad[8])
        dx = empty_cl(d.N, d.Cin, d.Din, d.Hin, d.Win, dev) if want_dx else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        if dx is not None or dw is not None:
            ws = workspace(lib.cfun_conv3d_bwd_fused_workspace_size(C.byref(d)), dev)
            _run("cfun_conv3d_bwd_fused_packed", C.byref(d), _ptr(xpack), xpack.numel(), _ptr(ypack), ypack.numel(), _ptr(w), _ptr(dx),
                 _ptr(dw), _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_BWD_WEIGHT, ALGO_AUTO) if _prof["on"] else "")
        dx2 = None
        if c2 > 0 and dx is not None:      # gradient of the concatenation -> its two sources
            da = empty_cl(d.N, c1, d.Din, d.Hin, d.Win, dev)
            dx2 = empty_cl(d.N, c2, d.Din, d.Hin, d.Win, dev)
            _run("cfun_split2_channels", _ptr(dx), c1, c2, _ptr(da), _ptr(dx2), d.N * d.Din * d.Hin * d.Win, _stream())
            dx = da
        return dx, dw, None, None, None, None, None, None, dx2


class UpNormConvInstNormActFn(Function):
    """The U-Net decoder's up block -- InstanceNorm3d -> LeakyReLU -> Upsample(x2, nearest) -> Conv3d(3^3) -> InstanceNorm3d ->
    LeakyReLU (reference mask_branch.py:91-103) -- as ONE autograd node: the first norm's apply pass writes the upsampled
    activation straight into the conv's operand pack (cfun_instnorm_up2_pack; the 8x larger fp32 tensor is never written or
    read), the conv epilogue accumulates the second norm's statistics, and the backward is ConvInstNormActFn's (the second
    norm writes the conv's dY pack) followed by the first norm's upsample-aware backward."""

    @staticmethod
    def forward(ctx, t, w, eps, slope):
        _require_cuda(t, w)
        t = to_cl(t)
        w = w.contiguous()
        N, C1, D, H, W = t.shape
        dev = t.device
        acc1 = getattr(t, "_cfun_in_stats", None)
        mean = torch.empty((N, C1), device=dev)
        rstd = torch.empty((N, C1), device=dev)
        if acc1 is not None:
            _run("cfun_instnorm_finalize", _ptr(acc1), N, D * H * W, C1, float(eps), _ptr(mean), _ptr(rstd), _stream())
        else:
            acc1 = torch.empty(2 * N * C1, dtype=torch.float64, device=dev)
            _run("cfun_instnorm_stats", _ptr(t), N, D * H * W, C1, float(eps), _ptr(acc1), _ptr(mean), _ptr(rstd), _stream())
        a1, b1 = _in_coeffs(mean, rstd, None, N, C1, eps)
        d = _conv_desc((N, C1, 2 * D, 2 * H, 2 * W), w.shape, 1, 1)
        nb = lib.cfun_conv3d_pack_bytes(C.byref(d))
        xpack = torch.empty(nb, dtype=torch.uint8, device=dev)
        G = ((C1 + 15) // 16) * 2
        _run("cfun_instnorm_up2_pack", _ptr(t), _ptr(a1), _ptr(b1), N, D, H, W, C1, float(slope), _ptr(xpack),
             C.c_void_p(xpack.data_ptr() + nb // 2), G, 1, _stream())
        Cc = d.Cout
        y = empty_cl(N, Cc, d.Dout, d.Hout, d.Wout, dev)
        ws = workspace(lib.cfun_conv3d_workspace_size(C.byref(d), PASS_FWD, ALGO_AUTO), dev)
        acc2 = torch.empty(2 * N * Cc, dtype=torch.float64, device=dev)
        _run("cfun_conv3d_fwd_stats_packed", C.byref(d), _ptr(xpack), nb, _ptr(w), _ptr(y), _ptr(acc2), _ptr(ws), ws.numel(), _stream(),
             tag=_conv_tag(d, PASS_FWD, ALGO_AUTO) if _prof["on"] else "")
        mean2 = torch.empty((N, Cc), device=dev)
        rstd2 = torch.empty((N, Cc), device=dev)
        _run("cfun_instnorm_finalize", _ptr(acc2), N, d.Dout * d.Hout * d.Wout, Cc, float(eps), _ptr(mean2), _ptr(rstd2), _stream())
        a2, b2 = _in_coeffs(mean2, rstd2, None, N, Cc, eps)
        z = empty_cl(N, Cc, d.Dout, d.Hout, d.Wout, dev)
        _run("cfun_affine_act_fwd", _ptr(y), _ptr(a2), _ptr(b2), Cc, None, _ptr(z), N, d.Dout, d.Hout, d.Wout, Cc, Cc, 0, 1, float(slope),
             _stream())
        ctx.save_for_backward(t, a1, b1, xpack, w, y, a2, b2)
        ctx.d, ctx.slope = d, float(slope)
        return z

    @staticmethod
    def backward(ctx, dz):
        t, a1, b1, xpack, w, y, a2, b2 = ctx.saved_tensors
        d, slope = ctx.d, ctx.slope
        N, Cc, D2, H2, W2 = d.N, d.Cout, d.Dout, d.Hout, d.Wout
        C1, D, H, W = d.Cin, D2 // 2, H2 // 2, W2 // 2
        dz = to_cl(dz)
        dev = dz.device
        # second norm + conv: as ConvInstNormActFn.backward
        g = empty_cl(N, Cc, D2, H2, W2, dev)
        acc2 = torch.empty(2 * N * Cc, dtype=torch.float64, device=dev)
        _run("cfun_affine_act_bwd", _ptr(y), _ptr(a2), _ptr(b2), Cc, None, _ptr(dz), _ptr(g), None, _ptr(acc2), N, D2, H2, W2, Cc, Cc, 0, 1,
             slope, _stream())
        Gy, Py = C.c_int(0), C.c_int(0)
        ybytes = lib.cfun_conv3d_dy_pack_geometry(C.byref(d), C.byref(Gy), C.byref(Py))
        ypack = torch.empty(ybytes, dtype=torch.uint8, device=dev)
        _run("cfun_instnorm_bwd_apply_pack", _ptr(y), _ptr(a2), _ptr(b2), _ptr(acc2), _ptr(g), N, D2, H2, W2, Cc, _ptr(ypack),
             C.c_void_p(ypack.data_ptr() + ybytes // 2), Gy.value, Py.value, _stream())
        del g
        need_t = ctx.needs_input_grad[0]
        du = empty_cl(N, C1, D2, H2, W2, dev) if need_t else None           # gradient of the (never materialised) upsampled activation
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        if du is not None or dw is not None:
            ws = workspace(lib.cfun_conv3d_bwd_fused_workspace_size(C.byref(d)), dev)
            _run("cfun_conv3d_bwd_fused_packed", C.byref(d), _ptr(xpack), xpack.numel(), _ptr(ypack), ypack.numel(), _ptr(w), _ptr(du),
                 _ptr(dw), _ptr(ws), ws.numel(), _stream(), tag=_conv_tag(d, PASS_BWD_WEIGHT, ALGO_AUTO) if _prof["on"] else "")
        dt = None
        if need_t:      # first norm: sums the 2x2x2 gradients, then the usual two passes
            dt = empty_cl(N, C1, D, H, W, dev)
            acc1 = torch.empty(2 * N * C1, dtype=torch.float64, device=dev)
            _run("cfun_affine_act_bwd", _ptr(t), _ptr(a1), _ptr(b1), C1, None, _ptr(du), _ptr(dt), None, _ptr(acc1), N, D, H, W, C1, C1, 0, 2,
                 slope, _stream())
            _run("cfun_instnorm_bwd_apply", _ptr(t), _ptr(a1), _ptr(b1), _ptr(acc1), _ptr(dt), N, D * H * W, C1, _stream())
        return dt, dw, None, None


def upnorm_conv_in_lrelu(t, w, b=None, eps=1e-5, slope=0.01):
    """instnorm_lrelu(conv3d(instnorm_lrelu(t, up=2), w, b, 1, 1)) -- the decoder's up block; one fused node where the conv has
    the fused tcgen05 backward and no bias, else the separate ops."""
    if b is None and _default_algo["algo"] == ALGO_AUTO and t.is_cuda and w.requires_grad and torch.is_grad_enabled() and \
            t.shape[1] % 4 == 0 and tuple(w.shape[2:]) == (3, 3, 3):
        N, C1, D, H, W = t.shape
        d = _conv_desc((N, C1, 2 * D, 2 * H, 2 * W), tuple(w.shape), 1, 1)
        if d.Cout % 4 == 0 and lib.cfun_conv3d_pack_bytes(C.byref(d)) and lib.cfun_conv3d_dy_pack_geometry(C.byref(d), None, None):
            return UpNormConvInstNormActFn.apply(t, w, eps, slope)
    return conv_in_lrelu(instnorm_lrelu(t, None, eps, slope, 2), w, b, 1, 1, None, eps, slope)


def conv_in_lrelu(x, w, b=None, stride=1, padding=0, drop=None, eps=1e-5, slope=0.01, up=1, x2=None):
    """instnorm_lrelu(conv3d(x, w, b), drop) -- as one fused node (ConvInstNormActFn) where the conv runs with the fused tcgen05
    backward and has no bias, else as the two separate ops (with the epilogue statistics where available).
    x2: the conv input is torch.cat((x, x2), dim=1); in the fused node the concatenation is never materialised."""
    if b is None and _default_algo["algo"] == ALGO_AUTO and x.is_cuda and w.requires_grad and torch.is_grad_enabled():
        cin = x.shape[1] + (0 if x2 is None else x2.shape[1])
        d = _conv_desc((x.shape[0], cin) + tuple(x.shape[2:]), tuple(w.shape), stride, padding)
        if d.Cout % 4 == 0 and lib.cfun_conv3d_pack_bytes(C.byref(d)) and lib.cfun_conv3d_dy_pack_geometry(C.byref(d), None, None):
            if x2 is None:
                return ConvInstNormActFn.apply(x, w, drop, stride, padding, eps, slope, up, None)
            if x2.shape[0] == x.shape[0] and x2.shape[2:] == x.shape[2:] and \
                    lib.cfun_conv3d_cat_supported(C.byref(d), int(x.shape[1]), int(x2.shape[1])):
                return ConvInstNormActFn.apply(x, w, drop, stride, padding, eps, slope, up, x2)
    if x2 is not None:
        x = cat_channels(x, x2)
    return instnorm_lrelu(conv3d(x, w, b, stride, padding, False, in_stats=True), drop, eps, slope, up)


class AddActInstNormFn(Function):
    """(leaky_relu(a + b), leaky_relu(InstanceNorm3d(a + b))) as one node: the U-Net's level-1 residual sum feeds both the skip
    connection and the norm (reference mask_branch.py:132-136).  Forward: one pass writes the sum, the skip activation and
    the norm statistics, one pass applies the norm; backward: the statistics pass, then ONE pass adds the norm's and the skip's
    gradients of the sum (unfused: add, lrelu, stats, apply forward; norm backward, lrelu backward, accumulate backward)."""

    @staticmethod
    def forward(ctx, a, b, eps, slope):
        _require_cuda(a, b)
        a, b = to_cl(a), to_cl(b)
        N, Cc, D, H, W = a.shape
        S = D * H * W
        dev = a.device
        s, c, y = (empty_cl(N, Cc, D, H, W, dev) for _ in range(3))
        acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=dev)
        mean = torch.empty((N, Cc), device=dev)
        rstd = torch.empty((N, Cc), device=dev)
        _run("cfun_add_act_stats", _ptr(a), _ptr(b), _ptr(s), _ptr(c), N, S, Cc, float(slope), float(eps), _ptr(acc), _ptr(mean),
             _ptr(rstd), _stream())
        ca, cb = _in_coeffs(mean, rstd, None, N, Cc, eps)
        _run("cfun_affine_act_fwd", _ptr(s), _ptr(ca), _ptr(cb), Cc, None, _ptr(y), N, D, H, W, Cc, Cc, 0, 1, float(slope), _stream())
        ctx.save_for_backward(s, ca, cb)
        ctx.slope = float(slope)
        return c, y

    @staticmethod
    def backward(ctx, dc, dy):
        s, ca, cb = ctx.saved_tensors
        N, Cc, D, H, W = s.shape
        dc, dy = to_cl(dc), to_cl(dy)
        ds = empty_cl(N, Cc, D, H, W, s.device)
        acc = torch.empty(2 * N * Cc, dtype=torch.float64, device=s.device)
        _run("cfun_affine_act_bwd", _ptr(s), _ptr(ca), _ptr(cb), Cc, None, _ptr(dy), _ptr(ds), None, _ptr(acc), N, D, H, W, Cc, Cc, 0, 1,
             ctx.slope, _stream())
        _run("cfun_instnorm_bwd_apply_extra", _ptr(s), _ptr(ca), _ptr(cb), _ptr(acc), _ptr(ds), N, D * H * W, Cc, _ptr(dc), ctx.slope,
             _stream())
        return ds, ds, None, None


def add_lrelu_instnorm(a, b, eps=1e-5, slope=0.01):
    """s = a + b -> (leaky_relu(s), instnorm_lrelu(s))"""
    return AddActInstNormFn.apply(a, b, eps, slope)


def instnorm_lrelu(x, drop=None, eps=1e-5, slope=0.01, up=1):
    return InstNormActFn.apply(x, drop, eps, slope, up, getattr(x, "_cfun_in_stats", None))


class MaxPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = to_cl(x)
        N, Cc, D, H, W = x.shape
        y = empty_cl(N, Cc, D // 2, H // 2, W // 2, x.device)
        _run("cfun_maxpool2_fwd", _ptr(x), _ptr(y), N, D, H, W, Cc, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, Cc, D, H, W = x.shape
        dy = to_cl(dy)
        dx = empty_cl(N, Cc, D, H, W, x.device)
        _run("cfun_maxpool2_bwd", _ptr(x), None, _ptr(dy), _ptr(dx), N, D, H, W, Cc, _stream())
        return dx


def maxpool2(x):
    return MaxPool2Fn.apply(x)


# --------------------------------------------------------------------------------------------------------
# RoI crop + resize
# --------------------------------------------------------------------------------------------------------
class RoiCropResizeFn(Function):
    @staticmethod
    def forward(ctx, f0, f1, boxes, level, pool, out_ncdhw):
        _require_cuda(f0, f1, boxes, level)
        f0 = to_cl(f0)
        f1 = to_cl(f1) if f1 is not None else None
        boxes = boxes.detach().contiguous().float()
        n = boxes.shape[0]
        Cc = f0.shape[1]
        pd, ph, pw = [int(p) for p in pool]
        if out_ncdhw:
            out = torch.empty((n, Cc, pd, ph, pw), device=f0.device)
        else:
            out = empty_cl(n, Cc, pd, ph, pw, f0.device)
        D1, H1, W1 = (f1.shape[2], f1.shape[3], f1.shape[4]) if f1 is not None else (0, 0, 0)
        _run("cfun_roi_crop_resize_fwd", _ptr(f0), f0.shape[2], f0.shape[3], f0.shape[4], _ptr(f1), D1, H1, W1, Cc,
             _ptr(boxes), _ptr(level), n, pd, ph, pw, _ptr(out), 1 if out_ncdhw else 0, _stream())
        ctx.save_for_backward(boxes, level)
        ctx.cfg = (tuple(f0.shape), tuple(f1.shape) if f1 is not None else None, (pd, ph, pw), out_ncdhw)
        return out

    @staticmethod
    def backward(ctx, dout):
        boxes, level = ctx.saved_tensors
        s0, s1, (pd, ph, pw), out_ncdhw = ctx.cfg
        n = boxes.shape[0]
        dout = dout.contiguous() if out_ncdhw else to_cl(dout)
        need0, need1 = ctx.needs_input_grad[0], (s1 is not None and ctx.needs_input_grad[1])
        if not (need0 or need1):
            return None, None, None, None, None, None
        df0 = zeros_cl(*s0, device=dout.device)
        df1 = zeros_cl(*s1, device=dout.device) if s1 is not None else None
        D1, H1, W1 = (s1[2], s1[3], s1[4]) if s1 is not None else (0, 0, 0)
        _run("cfun_roi_crop_resize_bwd", _ptr(df0), s0[2], s0[3], s0[4], _ptr(df1), D1, H1, W1, s0[1], _ptr(boxes),
             _ptr(level), n, pd, ph, pw, _ptr(dout), 1 if out_ncdhw else 0, _stream())
        return (df0 if need0 else None), (df1 if need1 else None), None, None, None, None


def roi_crop_resize(f0, f1, boxes, level, pool, out_ncdhw=False):
    return RoiCropResizeFn.apply(f0, f1, boxes, level, tuple(pool), out_ncdhw)


def roi_level(boxes):
    _require_cuda(boxes)
    boxes = boxes.detach().contiguous().float()
    lv = torch.empty(boxes.shape[0], dtype=torch.int32, device=boxes.device)
    _run("cfun_roi_level", _ptr(boxes), boxes.shape[0], _ptr(lv), _stream())
    return lv


# --------------------------------------------------------------------------------------------------------
# boxes
# --------------------------------------------------------------------------------------------------------
def sort_desc(scores):
    """indices of `scores` (1-D fp32) in (score descending, index ascending) order, int32"""
    _require_cuda(scores)
    scores = scores.detach().contiguous().float()
    n = scores.numel()
    order = torch.empty(n, dtype=torch.int32, device=scores.device)
    if n == 0:
        return order
    nb = lib.cfun_sort_workspace_size(n)
    ws = workspace(nb, scores.device)
    _run("cfun_sort_desc", _ptr(scores), n, _ptr(order), _ptr(ws), ws.numel(), _stream())
    return order


def decode_clip(anchors, deltas, scores, order, k, std6, window6, score_stride=1, score_offset=0):
    _require_cuda(anchors, deltas, scores, order)
    boxes = torch.empty((k, 6), device=anchors.device)
    sc = torch.empty(k, device=anchors.device) if scores is not None else None
    _run("cfun_decode_clip", _ptr(anchors.contiguous()), _ptr(deltas.detach().contiguous()),
         _ptr(scores.detach().contiguous()) if scores is not None else None, score_stride, score_offset, _ptr(order), k,
         f6(std6), f6(window6), _ptr(boxes), _ptr(sc), _stream())
    return boxes, sc


def nms3d(boxes_sorted, threshold, max_num):
    """boxes_sorted [n,6] already in descending score order.  Returns (keep int32[max_num], count int32[1]) on device."""
    _require_cuda(boxes_sorted)
    b = boxes_sorted.detach().contiguous().float()
    n = b.shape[0]
    keep = torch.zeros(max(max_num, 1), dtype=torch.int32, device=b.device)
    count = torch.zeros(1, dtype=torch.int32, device=b.device)
    nb = lib.cfun_nms_workspace_size(n)
    ws = workspace(nb, b.device)
    _run("cfun_nms3d", _ptr(b), n, float(threshold), int(max_num), _ptr(keep), _ptr(count), _ptr(ws), ws.numel(), _stream())
    return keep, count


def gather_boxes(rows, idx, count, max_rows, div6):
    out = torch.empty((max_rows, 6), device=rows.device)
    _run("cfun_gather_boxes", _ptr(rows.contiguous()), _ptr(idx), _ptr(count), max_rows, f6(div6), _ptr(out), _stream())
    return out


def iou_with_eps(box, boxes):
    """utils.compute_iou: [1,6] against [n,6] -> [1,n]"""
    _require_cuda(box, boxes)
    box = box.detach().contiguous().float()
    boxes = boxes.detach().contiguous().float()
    out = torch.empty((1, boxes.shape[0]), device=boxes.device)
    _run("cfun_iou3d_eps", _ptr(box), _ptr(boxes), boxes.shape[0], _ptr(out), _stream())
    return out


def bbox_overlaps3d(b1, b2):
    _require_cuda(b1, b2)
    b1 = b1.detach().contiguous().float()
    b2 = b2.detach().contiguous().float()
    out = torch.empty((b1.shape[0], b2.shape[0]), device=b1.device)
    _run("cfun_bbox_overlaps3d", _ptr(b1), b1.shape[0], _ptr(b2), b2.shape[0], _ptr(out), _stream())
    return out


def box_refinement(box, gt_box, std6=(1, 1, 1, 1, 1, 1)):
    _require_cuda(box, gt_box)
    box = box.detach().contiguous().float()
    gt_box = gt_box.detach().contiguous().float()
    out = torch.empty_like(box)
    _run("cfun_box_refinement", _ptr(box), _ptr(gt_box), box.shape[0], f6(std6), _ptr(out), _stream())
    return out


def roi_candidates(boxes_sorted, keep, count, div6, gt_boxes_norm, iou_threshold):
    """Device half A of detection_target_layer: -> (rois [max,6] normalised, assign int32 [max], pos_list, neg_list int32 [max],
    counts int32 [3] = {proposals, positives, negatives}); nothing is read back."""
    _require_cuda(boxes_sorted, keep, count, gt_boxes_norm)
    dev = boxes_sorted.device
    max_rows = keep.shape[0]
    gt = gt_boxes_norm.detach().contiguous().float()
    rois = torch.empty((max_rows, 6), device=dev)
    iou_max = torch.empty(max_rows, device=dev)
    ints = torch.empty((3, max_rows), dtype=torch.int32, device=dev)
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    _run("cfun_roi_candidates", _ptr(boxes_sorted.contiguous()), _ptr(keep), _ptr(count), max_rows, f6(div6), _ptr(gt), gt.shape[0],
         float(iou_threshold), _ptr(rois), _ptr(iou_max), _ptr(ints[0]), _ptr(ints[1]), _ptr(ints[2]), _ptr(counts), _stream())
    return rois, ints[0], ints[1], ints[2], counts


def roi_targets(rois, assign, pos_list, neg_list, perm, P, R, gt_boxes_norm, gt_class_ids, std6):
    """Device half B: perm int64 [R] on the device (positions in pos_list for rows < P, in neg_list after).
    -> (rois [R,6], class ids int64 [R], deltas [R,6])."""
    dev = rois.device
    gt = gt_boxes_norm.detach().contiguous().float()
    cls_in = gt_class_ids.to(torch.int32).contiguous()
    out = torch.empty((R, 6), device=dev)
    cls = torch.empty(R, dtype=torch.int64, device=dev)
    deltas = torch.empty((R, 6), device=dev)
    _run("cfun_roi_targets", _ptr(rois), _ptr(assign), _ptr(pos_list), _ptr(neg_list), _ptr(perm), int(P), int(R), _ptr(gt),
         _ptr(cls_in), f6(std6), _ptr(out), _ptr(cls), _ptr(deltas), _stream())
    return out, cls, deltas


def mask_target_crop(label_dhw, rois, ncls, mask_shape, onehot=True, index=True):
    """label_dhw int32 [D,H,W]; rois [P,6] normalised.  Returns (onehot float64 [P,ncls,*mask_shape] | None,
    class index int64 [P,*mask_shape] | None)."""
    _require_cuda(label_dhw, rois)
    label_dhw = label_dhw.contiguous()
    assert label_dhw.dtype == torch.int32
    rois = rois.detach().contiguous().float()
    P = rois.shape[0]
    md, mh, mw = [int(m) for m in mask_shape]
    oh = torch.empty((P, ncls, md, mh, mw), dtype=torch.float64, device=rois.device) if onehot else None
    ix = torch.empty((P, md, mh, mw), dtype=torch.int64, device=rois.device) if index else None
    D, H, W = label_dhw.shape
    _run("cfun_mask_target_crop", _ptr(label_dhw), D, H, W, _ptr(rois), P, ncls, md, mh, mw, _ptr(oh), _ptr(ix), _stream())
    return oh, ix


# --------------------------------------------------------------------------------------------------------
# Sobel edge loss
# --------------------------------------------------------------------------------------------------------
class SobelEdgeLossFn(Function):
    @staticmethod
    def forward(ctx, pred, tgt_index, mode):
        _require_cuda(pred, tgt_index)
        pred = to_cl(pred)
        P, ncls, Md, Mh, Mw = pred.shape
        tgt_index = tgt_index.contiguous()
        assert tgt_index.dtype == torch.int64 and tuple(tgt_index.shape) == (P, Md, Mh, Mw)
        loss = torch.empty(1, device=pred.device)
        ws = workspace(4096, pred.device)
        _run("cfun_sobel_edge_loss_fwd", _ptr(pred), _ptr(tgt_index), P, Md, Mh, Mw, ncls, mode, _ptr(loss), _ptr(ws), ws.numel(),
             _stream())
        ctx.save_for_backward(pred, tgt_index)
        ctx.mode = mode
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pred, tgt_index = ctx.saved_tensors
        P, ncls, Md, Mh, Mw = pred.shape
        g = dloss.reshape(1).contiguous().float()
        dpred = empty_cl(P, ncls, Md, Mh, Mw, pred.device)
        nb = lib.cfun_sobel_edge_workspace_size(P, Md, Mh, Mw, ncls, ctx.mode)
        ws = workspace(nb, pred.device)
        _run("cfun_sobel_edge_loss_bwd", _ptr(pred), _ptr(tgt_index), P, Md, Mh, Mw, ncls, ctx.mode, _ptr(g), _ptr(dpred), _ptr(ws),
             ws.numel(), _stream())
        return dpred, None, None


SOBEL_MAGNITUDE, SOBEL_RAW = 0, 1


def sobel_edge_loss(pred_probs, tgt_index, mode=SOBEL_MAGNITUDE):
    """pred_probs [P,ncls,d,h,w], tgt_index int64 [P,d,h,w].  mode SOBEL_MAGNITUDE: reference model.py:938-981 (heart);
    SOBEL_RAW: LiTS_2017/model.py:943-981."""
    return SobelEdgeLossFn.apply(pred_probs, tgt_index, mode)


class MaskCEFn(Function):
    """nn.CrossEntropyLoss(weight) between mask logits [P,ncls,d,h,w] and class-index targets [P,d,h,w] (reference
    model.py:909-935, LiTS_2017/model.py:926): one fused pass forward, one backward."""

    @staticmethod
    def forward(ctx, logits, target, weight):
        _require_cuda(logits, target, weight)
        logits = to_cl(logits)
        P, ncls = logits.shape[0], logits.shape[1]
        target = target.contiguous()
        assert target.dtype == torch.int64 and target.numel() * ncls == logits.numel()
        if weight is not None:
            weight = weight.contiguous().float()
        acc = torch.empty(2, dtype=torch.float64, device=logits.device)
        loss = torch.empty((), device=logits.device)
        _run("cfun_mask_ce_fwd", _ptr(logits), _ptr(target), target.numel(), ncls, _ptr(weight), _ptr(acc), _ptr(loss), _stream())
        ctx.save_for_backward(logits, target, weight, acc)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, target, weight, acc = ctx.saved_tensors
        g = dloss.reshape(1).contiguous().float()
        P, ncls, D, H, W = logits.shape
        dl = empty_cl(P, ncls, D, H, W, logits.device)
        _run("cfun_mask_ce_bwd", _ptr(logits), _ptr(target), target.numel(), ncls, _ptr(weight), _ptr(acc), _ptr(g), _ptr(dl), _stream())
        return dl, None, None


def mask_cross_entropy(logits, target_index, weight=None):
    return MaskCEFn.apply(logits, target_index, weight)


# --------------------------------------------------------------------------------------------------------
# optimizer tail / input molding
# --------------------------------------------------------------------------------------------------------
def sumsq_into(flat_grad, acc):
    _run("cfun_sumsq", _ptr(flat_grad), flat_grad.numel(), _ptr(acc), _stream())


def sgd_clip_step(p, g, mom, wd_mask, sumsq, max_norm, lr, momentum, weight_decay):
    _run("cfun_sgd_clip_step", _ptr(p), _ptr(g), _ptr(mom), _ptr(wd_mask), p.numel(), _ptr(sumsq), float(max_norm),
         float(lr), float(momentum), float(weight_decay), 0, _stream())


def resize_linear3d(vol_hwd, out_shape):
    """order-1 resize of a raw scan [H,W,D] (int16 or float32, on device) to out_shape, reference utils.resize_image
    'self' mode semantics (utils.py:389-393); same dtype out"""
    _require_cuda(vol_hwd)
    assert vol_hwd.dim() == 3 and vol_hwd.dtype in (torch.int16, torch.float32)
    vol_hwd = vol_hwd.contiguous()
    H, W, D = vol_hwd.shape
    H2, W2, D2 = [int(v) for v in out_shape]
    out = torch.empty((H2, W2, D2), dtype=vol_hwd.dtype, device=vol_hwd.device)
    _run("cfun_resize_linear3d", _ptr(vol_hwd), H, W, D, _ptr(out), H2, W2, D2, 0 if vol_hwd.dtype == torch.int16 else 1, _stream())
    return out


def unmold_mask_argmax(mask_cdhw, box6, image_dhw):
    """class-probability crop of one detection [ncls,d,h,w] (device, any dense layout) -> uint8 class-id volume [H,W,D]:
    utils.unmold_mask + argmax (reference utils.py:443-460, model.py:1851-1853) fused"""
    _require_cuda(mask_cdhw)
    m = mask_cdhw.float()
    ncls, md, mh, mw = m.shape
    vs = m.stride(3)
    if not (m.stride(2) == mw * vs and m.stride(1) == mh * mw * vs):
        m = m.contiguous()
        vs = 1
    D, H, W = [int(v) for v in image_dhw]
    out = torch.empty((H, W, D), dtype=torch.uint8, device=m.device)
    box = (C.c_int * 6)(*[int(v) for v in box6])
    _run("cfun_unmold_mask_argmax", _ptr(m), ncls, md, mh, mw, m.stride(0), vs, box, D, H, W, _ptr(out), _stream())
    return out


def mold_volume_i16(vol_hwd):
    """int16 [H,W,D] CT volume on device -> molded fp32 [1,1,D,H,W] (model.mold_image + the HWD->DHW transpose)."""
    _require_cuda(vol_hwd)
    assert vol_hwd.dtype == torch.int16 and vol_hwd.dim() == 3
    vol_hwd = vol_hwd.contiguous()
    H, W, D = vol_hwd.shape
    out = torch.empty((1, 1, D, H, W), device=vol_hwd.device)
    acc = torch.empty(2, dtype=torch.float64, device=vol_hwd.device)
    _run("cfun_mold_volume_i16", _ptr(vol_hwd), H, W, D, _ptr(acc), _ptr(out), _stream())
    return out
