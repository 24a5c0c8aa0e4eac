"""Thin torch-tensor wrappers over the C ABI (include/ghnd_b200.h).

torch is used for device memory and streams only; every function here enqueues hand-written
sm_100a kernels from libghnd_b200.so on the current CUDA stream and raises GhndError on failure.
There is no CPU path.
"""
import ctypes
from ctypes import byref, c_float, c_void_p

import torch

from . import _lib
from ._lib import BF16, F16, call, fmt_of, ptr, stream_ptr

# kernel-launch accounting (bench.py reports it as gpu_launches)
_launches = 0


def launches():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.GhndError("ghnd ops need CUDA tensors (no CPU fallback); got %s" % t.device)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# quantizer
# ------------------------------------------------------------------------------------------------
def quantize_ws(n, device):
    """Zeroed workspace of ghnd_quantize_u8 (its grid-barrier words must be zero on entry; the kernel
    re-arms them on exit, so one workspace serves any number of calls on one stream)."""
    return torch.zeros(_lib.load().ghnd_quantize_u8_workspace_bytes(n), dtype=torch.uint8, device=device)


def quantize_u8(x, num_bits=8, scale_mode=_lib.QSCALE_DIV, q=None, qparams=None, ws=None):
    """-> (q uint8 like x, qparams: 16-byte device tensor {scale f32, zp i32, min f32, max f32})."""
    _need_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    if q is None:
        q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if qparams is None:
        qparams = torch.empty(4, dtype=torch.int32, device=x.device)
    n = x.numel()
    if ws is None:
        ws = quantize_ws(n, x.device)
    call("ghnd_quantize_u8", ptr(x), n, num_bits, scale_mode, ptr(q), ptr(qparams), ptr(ws), ws.numel(), stream_ptr())
    _count(1)
    return q, qparams


def quantize_u8_minmax(x, minmax, n_pairs, num_bits=8, scale_mode=_lib.QSCALE_DIV, q=None, qparams=None):
    """Single-pass quantizer for a tensor whose (min, max) partial pairs were published by its
    producer (conv_narrow_out(..., minmax=...)).  Same bytes as quantize_u8."""
    _need_cuda(x, minmax)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if q is None:
        q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if qparams is None:
        qparams = torch.empty(4, dtype=torch.int32, device=x.device)
    call("ghnd_quantize_u8_minmax", ptr(x), x.numel(), num_bits, scale_mode, ptr(minmax), int(n_pairs),
         ptr(q), ptr(qparams), stream_ptr())
    _count(1)
    return q, qparams


def dequantize_u8(q, qparams, out=None):
    _need_cuda(q, qparams)
    q = q.contiguous()
    if out is None:
        out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    if q.numel():
        call("ghnd_dequantize_u8", ptr(q), q.numel(), ptr(qparams), ptr(out), stream_ptr())
        _count(1)
    return out


def make_qparams(scale, zero_point, device):
    qp = torch.zeros(4, dtype=torch.int32, device=device)
    qp[0:1].view(torch.float32).copy_(torch.as_tensor([float(scale)], dtype=torch.float32))
    qp[1] = int(zero_point)
    return qp


# ------------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------------
def sse_fwd_bwd(levels, grad_dtype=torch.bfloat16, loss_out=None, workspace=None):
    """levels: list of (teacher, student, grad_or_None, factor, relu_mask) with 16-bit tensors of
    identical shape/dtype.  Returns float32 tensor [1+n]: total, then factor*SSE per level."""
    n = len(levels)
    arr = (_lib.SseLevel * n)()
    dev = levels[0][0].device
    in_fmt = fmt_of(levels[0][0].dtype)
    for i, (t, s, g, factor, mask) in enumerate(levels):
        _need_cuda(t, s, g)
        assert t.shape == s.shape and t.dtype == s.dtype and t.is_contiguous() and s.is_contiguous()
        arr[i].teacher = t.data_ptr()
        arr[i].student = s.data_ptr()
        arr[i].grad = g.data_ptr() if g is not None else None
        arr[i].n = t.numel()
        arr[i].factor = float(factor)
        arr[i].relu_mask = int(mask)
    if loss_out is None:
        loss_out = torch.empty(1 + n, dtype=torch.float32, device=dev)
    wsz = _lib.load().ghnd_sse_workspace_bytes()
    if workspace is None:
        workspace = _ws(wsz, dev)
    call("ghnd_sse_fwd_bwd", arr, n, in_fmt, fmt_of(grad_dtype), ptr(loss_out), ptr(workspace), wsz,
         stream_ptr())
    _count(2)
    return loss_out


# ------------------------------------------------------------------------------------------------
# layout boundary
# ------------------------------------------------------------------------------------------------
def to_nhwc16(x, dtype=torch.float16):
    """NCHW fp32 -> NHWC 16-bit tensor of shape [N,H,W,C]."""
    _need_cuda(x)
    x = x.float().contiguous()
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=dtype, device=x.device)
    call("ghnd_nchw_f32_to_nhwc16", ptr(x), ptr(y), fmt_of(dtype), n, c, h, w, stream_ptr())
    _count()
    return y


def to_nhwc16_into(x, y):
    """NCHW fp32 -> existing NHWC 16-bit buffer y [N,H,W,C]."""
    _need_cuda(x, y)
    x = x.float().contiguous()
    n, c, h, w = x.shape
    assert tuple(y.shape) == (n, h, w, c), (tuple(y.shape), (n, h, w, c))
    call("ghnd_nchw_f32_to_nhwc16", ptr(x), ptr(y), fmt_of(y.dtype), n, c, h, w, stream_ptr())
    _count()
    return y


def to_nchw_f32(y):
    """NHWC 16-bit [N,H,W,C] -> NCHW fp32."""
    _need_cuda(y)
    y = y.contiguous()
    n, h, w, c = y.shape
    x = torch.empty((n, c, h, w), dtype=torch.float32, device=y.device)
    call("ghnd_nhwc16_to_nchw_f32", ptr(y), fmt_of(y.dtype), ptr(x), n, c, h, w, stream_ptr())
    _count()
    return x


# ------------------------------------------------------------------------------------------------
# plans (wide convs)
# ------------------------------------------------------------------------------------------------
class ConvPlan(object):
    """ghnd_conv_plan: tcgen05 implicit-GEMM conv forward / dgrad bound to fixed buffers."""

    def __init__(self, kind, N, H, W, C, K, R, S, stride, pad, src, weights, dst, bias=None,
                 residual=None, relu=False, mask=None, accumulate=False, stats=None, stats_mode=0,
                 mask_stats_only=False, stats_zeroed=False):
        d = _lib.ConvDesc()
        d.kind = kind
        d.N, d.H, d.W, d.C, d.K, d.R, d.S, d.stride, d.pad = N, H, W, C, K, R, S, stride, pad
        d.src, d.src_fmt = src.data_ptr(), fmt_of(src.dtype)
        d.weights, d.w_fmt = weights.data_ptr(), fmt_of(weights.dtype)
        d.dst, d.dst_fmt = dst.data_ptr(), fmt_of(dst.dtype)
        d.bias = bias.data_ptr() if bias is not None else None
        d.residual = residual.data_ptr() if residual is not None else None
        d.res_fmt = fmt_of(residual.dtype) if residual is not None else 0
        d.relu = int(bool(relu))
        d.mask = mask.data_ptr() if mask is not None else None
        d.mask_fmt = fmt_of(mask.dtype) if mask is not None else 0
        d.accumulate = int(bool(accumulate))
        if stats is not None:
            assert stats.dtype == torch.float64 and stats.numel() >= 2 * (K if kind == _lib.CONV_FWD else C)
        d.stats = stats.data_ptr() if stats is not None else None
        d.stats_mode = int(stats_mode) | (_lib.SUMS_ZEROED if (stats_zeroed and stats is not None) else 0)
        d.mask_stats_only = int(bool(mask_stats_only))
        self._keep = (src, weights, dst, bias, residual, mask, stats)  # buffers must outlive the plan
        self.desc = "%s N%d %dx%d C%d K%d %dx%d s%d p%d%s%s%s%s" % (
            "fwd" if kind == _lib.CONV_FWD else "dgrad", N, H, W, C, K, R, S, stride, pad,
            " +res" if residual is not None else "", " relu" if relu else "",
            " mask" if mask is not None else "", " acc" if accumulate else "")
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        self.flops = 2.0 * N * Ho * Wo * K * C * R * S
        self._h = c_void_p()
        call("ghnd_conv_plan_create", byref(d), byref(self._h))
        self.n_launches = _lib.load().ghnd_conv_plan_launches(self._h)

    def run(self, stream=None):
        try:
            call("ghnd_conv_plan_run", self._h, stream_ptr(stream))
        except _lib.GhndError as e:
            raise _lib.GhndError("%s [conv plan: %s]" % (e, self.desc))
        _count(self.n_launches)

    def run_range(self, first, count, stream=None):
        """Only launches [first, first+count): they write disjoint parts of dst, so the caller may
        put them on different streams (parity classes of a stride-2 dgrad)."""
        call("ghnd_conv_plan_run_range", self._h, int(first), int(count), stream_ptr(stream))
        _count(count)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ghnd_conv_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class WgradPlan(object):
    def __init__(self, N, H, W, C, K, R, S, pad, x, dy, dw):
        d = _lib.WgradDesc()
        d.N, d.H, d.W, d.C, d.K, d.R, d.S, d.pad = N, H, W, C, K, R, S, pad
        d.x, d.x_fmt = x.data_ptr(), fmt_of(x.dtype)
        d.dy, d.dy_fmt = dy.data_ptr(), fmt_of(dy.dtype)
        assert dw.dtype == torch.float32 and dw.numel() == K * R * S * C
        d.dw = dw.data_ptr()
        self._keep = (x, dy, dw)
        self.desc = "wgrad N%d %dx%d C%d K%d %dx%d p%d" % (N, H, W, C, K, R, S, pad)
        self.flops = 2.0 * N * (H + 2 * pad - R + 1) * (W + 2 * pad - S + 1) * K * C * R * S
        self._h = c_void_p()
        call("ghnd_wgrad_plan_create", byref(d), byref(self._h))

    def run(self, stream=None):
        call("ghnd_wgrad_plan_run", self._h, stream_ptr(stream))
        _count(1)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ghnd_wgrad_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class StemPlan(object):
    """conv1 7x7 s2 + folded FrozenBN + ReLU.  y [N,Hp/2,Wp/2,K] with K = 64 (one stem) or 128 (two
    stems over the same packed image as one GEMM: weights [K][7][32], bias [K])."""

    def __init__(self, x_packed, w_packed, bias, y, N, Hp, Wp):
        K = y.shape[-1]
        assert w_packed.shape[0] == K and bias.numel() == K
        self._keep = (x_packed, w_packed, bias, y)
        self.desc = "stem N%d %dx%d" % (N, Hp, Wp) + ("" if K == 64 else " K%d" % K)
        self.n_launches = 4
        self.flops = 2.0 * N * (Hp // 2) * (Wp // 2) * K * 147
        self._h = c_void_p()
        call("ghnd_stem_conv_plan_create_k", ptr(x_packed), fmt_of(x_packed.dtype), ptr(w_packed),
             fmt_of(w_packed.dtype), ptr(bias), ptr(y), fmt_of(y.dtype), N, Hp, Wp, K, byref(self._h))

    def run(self, stream=None):
        call("ghnd_stem_conv_plan_run", self._h, stream_ptr(stream))
        _count(4)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ghnd_stem_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class StemPoolPlan(object):
    """conv1 + folded FrozenBN + ReLU + MaxPool 3x3 s2 p1 in one kernel: the conv output stays in shared
    memory.  ys: one pooled [N,Ho,Wo,64] tensor per model (1 or 2 models over the same packed image,
    weights [64*m][7][32], bias [64*m]); argmaxes: matching uint8 tensors or None entries."""

    def __init__(self, x_packed, w_packed, bias, ys, argmaxes, N, Hp, Wp):
        m = len(ys)
        assert w_packed.shape[0] == 64 * m and bias.numel() == 64 * m
        argmaxes = list(argmaxes) if argmaxes is not None else [None] * m
        ho, wo = (Hp // 2 + 1) // 2, (Wp // 2 + 1) // 2
        for y, a in zip(ys, argmaxes):
            assert tuple(y.shape) == (N, ho, wo, 64) and y.is_contiguous() and y.dtype == x_packed.dtype
            assert a is None or (tuple(a.shape) == (N, ho, wo, 64) and a.dtype == torch.uint8 and a.is_contiguous())
        self._keep = (x_packed, w_packed, bias, list(ys), argmaxes)
        self.desc = "stem+pool N%d %dx%d" % (N, Hp, Wp) + ("" if m == 1 else " x%d" % m)
        self.n_launches = 1
        self.flops = 2.0 * N * (Hp // 2) * (Wp // 2) * 64 * m * 147
        yp = (c_void_p * m)(*[ptr(y) for y in ys])
        ap = (c_void_p * m)(*[ptr(a) for a in argmaxes])
        self._h = c_void_p()
        call("ghnd_stem_pool_plan_create", ptr(x_packed), fmt_of(x_packed.dtype), ptr(w_packed),
             fmt_of(w_packed.dtype), ptr(bias), m, yp, ap, fmt_of(x_packed.dtype), N, Hp, Wp, byref(self._h))

    def run(self, stream=None):
        call("ghnd_stem_pool_plan_run", self._h, stream_ptr(stream))
        _count()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ghnd_stem_pool_plan_destroy(h)
            except Exception:
                pass
            self._h = None


# ------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------
def pack_weight(w, scale=None, transpose=False, dtype=torch.float16, out=None):
    """OIHW fp32 (x per-O scale) -> [O][R][S][I] (or [I][R][S][O] when transpose) 16-bit."""
    _need_cuda(w)
    w = w.detach().float().contiguous()
    o, i, r, s = w.shape
    if out is None:
        out = torch.empty((i, r, s, o) if transpose else (o, r, s, i), dtype=dtype, device=w.device)
    call("ghnd_pack_weight", ptr(w), ptr(scale), o, i, r, s, int(transpose), ptr(out), fmt_of(out.dtype),
         stream_ptr())
    _count()
    return out


def pack_weights(items):
    """items: list of (w OIHW fp32, scale or None, transpose, out) -- the arguments of pack_weight -- repacked by
    ONE launch per GHND_PACK_MAX tensors (ghnd_pack_weights)."""
    for k0 in range(0, len(items), _lib.PACK_MAX):
        chunk = items[k0:k0 + _lib.PACK_MAX]
        descs = (_lib.PackWeightDesc * len(chunk))()
        keep = []
        for d, (w, scale, transpose, out) in zip(descs, chunk):
            _need_cuda(w)
            wf = w.detach()
            if wf.dtype != torch.float32 or not wf.is_contiguous():
                wf = wf.float().contiguous()
            keep.append(wf)
            o, i, r, s = wf.shape
            assert out.numel() == wf.numel()
            d.w_oihw, d.scale_o = wf.data_ptr(), (scale.data_ptr() if scale is not None else None)
            d.O, d.I, d.R, d.S, d.transpose = o, i, r, s, int(bool(transpose))
            d.dst, d.dst_fmt = out.data_ptr(), fmt_of(out.dtype)
        call("ghnd_pack_weights", descs, len(chunk), stream_ptr())
        _count()


def unpack_wgrad(dw, out, alpha=1.0):
    """[O][R][S][I] fp32 -> OIHW fp32 into `out`."""
    o, i, r, s = out.shape
    call("ghnd_unpack_wgrad", ptr(dw), ptr(out), o, i, r, s, float(alpha), stream_ptr())
    _count()
    return out


def stem_pack_weight(w, scale=None, dtype=torch.float16, out=None):
    w = w.detach().float().contiguous()
    if out is None:
        out = torch.empty((64, 7, 32), dtype=dtype, device=w.device)
    call("ghnd_stem_pack_weight", ptr(w), ptr(scale), ptr(out), fmt_of(out.dtype), stream_ptr())
    _count()
    return out


class ScaledImage:
    """An image that GeneralizedRCNNTransform.resize (src/models/org/rcnn.py:29-45) would resample by
    `scale`: the source tensor plus the output geometry; the bilinear resample itself happens inside
    the stem pack kernel (ghnd_stem_pack_image_resized).  `.shape` is the shape AFTER resizing, so
    callers that size the batch from `img.shape` work on either kind."""

    def __init__(self, src, scale):
        import math
        self.src = src
        self.scale = float(scale)
        h, w = src.shape[-2:]
        # ATen upsample output size: floor(double(in) * scale_factor)
        self.shape = torch.Size((src.shape[0], int(math.floor(float(h) * self.scale)),
                                 int(math.floor(float(w) * self.scale))))
        if self.shape[1] < 1 or self.shape[2] < 1:
            raise ValueError("scale_factor %g collapses a %dx%d image" % (self.scale, h, w))

    @property
    def is_cuda(self):
        return self.src.is_cuda

    @property
    def device(self):
        return self.src.device


def stem_pack_image(img, dst, n_index, Hp, Wp, mean, std):
    """img [3,H,W] fp32 in [0,1] (or a ScaledImage) -> dst[n_index] of the packed batch
    [N][Hp+6][Wp+8][4]."""
    m = (c_float * 3)(*[float(v) for v in mean])
    s = (c_float * 3)(*[float(v) for v in std])
    if isinstance(img, ScaledImage):
        src = img.src
        _need_cuda(src, dst)
        src = src.float().contiguous()
        rs = 1.0 / img.scale  # ATen: static_cast<float>(1.0 / scale_factor), same for both axes
        call("ghnd_stem_pack_image_resized", ptr(src), src.shape[1], src.shape[2], img.shape[1],
             img.shape[2], c_float(rs), c_float(rs), m, s, ptr(dst), fmt_of(dst.dtype), n_index, Hp, Wp,
             stream_ptr())
        _count()
        return
    _need_cuda(img, dst)
    img = img.float().contiguous()
    call("ghnd_stem_pack_image", ptr(img), img.shape[1], img.shape[2], m, s, ptr(dst), fmt_of(dst.dtype),
         n_index, Hp, Wp, stream_ptr())
    _count()


def stem_pack_images(imgs, dst, Hp, Wp, mean, std, n_index0=0):
    """All images of a batch (fp32 [3,H,W] in [0,1], or ScaledImage) -> dst[n_index0 ...] in ONE launch."""
    n = len(imgs)
    m = (c_float * 3)(*[float(v) for v in mean])
    s = (c_float * 3)(*[float(v) for v in std])
    srcs, keep = [], []
    H, W, Ho, Wo, rs = (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)(), (c_float * n)()
    for k, img in enumerate(imgs):
        src = img.src if isinstance(img, ScaledImage) else img
        _need_cuda(src, dst)
        src = src.float().contiguous()
        keep.append(src)
        srcs.append(ptr(src))
        H[k], W[k] = src.shape[1], src.shape[2]
        Ho[k], Wo[k] = img.shape[1], img.shape[2]
        # ATen: static_cast<float>(1.0 / scale_factor), same for both axes; 0 = no resize
        rs[k] = 1.0 / img.scale if isinstance(img, ScaledImage) else 0.0
    arr = (c_void_p * n)(*srcs)
    call("ghnd_stem_pack_images", arr, H, W, Ho, Wo, rs, n, m, s, ptr(dst), fmt_of(dst.dtype), n_index0, Hp, Wp,
         stream_ptr())
    _count((n + 15) // 16)


# ------------------------------------------------------------------------------------------------
# narrow convs (planar fp32 bottleneck side <-> NHWC 16-bit wide side)
# ------------------------------------------------------------------------------------------------
def _narrow_ws(c, k, r, s, device):
    n = _lib.load().ghnd_conv_narrow_workspace_bytes(c, k, r, s)
    return _ws(n, device), n


MINMAX_CAPACITY = 1024  # (min, max) pairs a producer may publish (>= 2 x SM count)


def conv_narrow_out(x, w, pad, y=None, ws=None, minmax=None):
    """x NHWC16 [N,H,W,C], w OIHW fp32 [K,C,R,S] -> y planar fp32 [N,K,Ho,Wo].
    minmax: optional fp32 [2*MINMAX_CAPACITY] buffer; the kernel then also writes per-CTA (min, max)
    pairs of y into it and the call returns (y, number of pairs) -- input of quantize_u8_minmax."""
    n, h, wd, c = x.shape
    k, _, r, s = w.shape
    ho, wo = h + 2 * pad - r + 1, wd + 2 * pad - s + 1
    if y is None:
        y = torch.empty((n, k, ho, wo), dtype=torch.float32, device=x.device)
    if ws is None:
        ws = _narrow_ws(c, k, r, s, x.device)
    if minmax is not None:
        assert minmax.dtype == torch.float32 and minmax.numel() >= 2
        n_pairs = ctypes.c_int(0)
        call("ghnd_conv_narrow_out_minmax", ptr(x), fmt_of(x.dtype), ptr(w), ptr(y), n, h, wd, c, k, r, s,
             pad, ptr(ws[0]), ws[1], ptr(minmax), minmax.numel() // 2, byref(n_pairs), stream_ptr())
        _count(1)
        return y, n_pairs.value
    call("ghnd_conv_narrow_out", ptr(x), fmt_of(x.dtype), ptr(w), ptr(y), n, h, wd, c, k, r, s, pad,
         ptr(ws[0]), ws[1], stream_ptr())
    _count(1)
    return y


def conv_narrow_in(x, w, pad, pre=None, pre_relu=False, flip=False, dtype=torch.float16, y=None, ws=None):
    """flip=False: x planar fp32 [N,C,H,W], w [K,C,R,S] -> y NHWC16 [N,Ho,Wo,K].
    flip=True : x = dy planar [N,C,H,W] of a narrow-out conv with weight w [C,K,R,S] -> dx NHWC16."""
    n, c, h, wd = x.shape
    if not flip:
        k, _, r, s = w.shape
        ho, wo = h + 2 * pad - r + 1, wd + 2 * pad - s + 1
    else:
        _, k, r, s = w.shape
        ho, wo = h - 2 * pad + r - 1, wd - 2 * pad + s - 1
    if y is None:
        y = torch.empty((n, ho, wo, k), dtype=dtype, device=x.device)
    if ws is None:
        ws = _narrow_ws(c, k, r, s, x.device)
    call("ghnd_conv_narrow_in", ptr(x), ptr(pre), int(bool(pre_relu)), ptr(w), int(bool(flip)), ptr(y),
         fmt_of(y.dtype), n, h, wd, c, k, r, s, pad, ptr(ws[0]), ws[1], stream_ptr())
    _count(2)
    return y


def conv_narrow_out_dgrad(dy, w, pad, H, W, dx=None, ws=None):
    """dgrad of a narrow-in conv (x planar [N,C,H,W] -> y NHWC [N,Ho,Wo,K], w [K,C,R,S]):
    dy NHWC16 -> dx planar fp32 [N,C,H,W]."""
    n = dy.shape[0]
    k, c, r, s = w.shape
    if dx is None:
        dx = torch.empty((n, c, H, W), dtype=torch.float32, device=dy.device)
    if ws is None:
        ws = _narrow_ws(c, k, r, s, dy.device)
    call("ghnd_conv_narrow_out_dgrad", ptr(dy), fmt_of(dy.dtype), ptr(w), ptr(dx), n, H, W, c, k, r, s, pad,
         ptr(ws[0]), ws[1], stream_ptr())
    _count(1)
    return dx


def wgrad_narrow(a, b, dw, a_is_output, R, S, pad, pre=None, pre_relu=False, ws=None):
    """a planar fp32 [N,Ca,Ha,Wa], b NHWC16 [N,Hb,Wb,Cb] -> dw OIHW fp32 (see ghnd_b200.h)."""
    n, ca, ha, wa = a.shape
    _, hb, wb, cb = b.shape
    nbytes = _lib.load().ghnd_wgrad_narrow_workspace_bytes(ca, cb, R, S)
    if ws is None:
        ws = _ws(nbytes, a.device)
    call("ghnd_wgrad_narrow", ptr(a), ptr(pre), int(bool(pre_relu)), ptr(b), fmt_of(b.dtype), ptr(dw),
         int(bool(a_is_output)), n, ha, wa, ca, hb, wb, cb, R, S, pad, ptr(ws), nbytes, stream_ptr())
    _count(1 + (ca + 2) // 3)
    return dw


# ------------------------------------------------------------------------------------------------
# stem helpers
# ------------------------------------------------------------------------------------------------
def maxpool3x3s2(x, y=None, argmax=None, channels=None, channel_offset=0):
    """channels / channel_offset: pool only channels [offset, offset+channels) of x (one half of the
    two-stem conv output); y / argmax are compact `channels`-wide tensors."""
    n, h, w, xc = x.shape
    c = xc if channels is None else channels
    ho, wo = (h + 1) // 2, (w + 1) // 2
    if y is None:
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
    call("ghnd_maxpool3x3s2_strided", ptr(x), xc, channel_offset, ptr(y), ptr(argmax), fmt_of(x.dtype), n, h, w,
         c, stream_ptr())
    _count()
    return y


def maxpool3x3s2_bwd(x, argmax, dy, dx, channel_offset=0):
    """x may be wider than dx (two-stem conv output): its channels [offset, offset+C) are used.  x only
    supplies geometry (the ReLU mask travels in the argmax codes); None = the geometry of dx, for the
    fused stem+pool path where the pool's input never exists."""
    if x is None:
        x, channel_offset = dx, 0
    n, h, w, xc = x.shape
    c = dx.shape[-1]
    call("ghnd_maxpool3x3s2_bwd_strided", ptr(x), fmt_of(x.dtype), xc, channel_offset, ptr(argmax), ptr(dy),
         fmt_of(dy.dtype), ptr(dx), fmt_of(dx.dtype), n, h, w, c, stream_ptr())
    _count()
    return dx


def stem_wgrad(x_packed, g, scale, dw, N, Hp, Wp, ws=None):
    nbytes = _lib.load().ghnd_stem_wgrad_workspace_bytes()
    if ws is None:
        ws = _ws(nbytes, g.device)
    call("ghnd_stem_wgrad", ptr(x_packed), fmt_of(x_packed.dtype), ptr(g), fmt_of(g.dtype), ptr(scale),
         ptr(dw), N, Hp, Wp, ptr(ws), nbytes, stream_ptr())
    _count(2)
    return dw


class StemWgradPlan(object):
    """conv1 weight gradient on tcgen05; x_packed and g must share one 16-bit dtype."""

    def __init__(self, x_packed, g, scale, dw, N, Hp, Wp, ws=None):
        assert x_packed.dtype == g.dtype, "stem wgrad: packed image and gradient must share one 16-bit dtype"
        nbytes = _lib.load().ghnd_stem_wgrad_workspace_bytes()
        if ws is None:
            ws = _ws(nbytes, g.device)
        self._keep = (x_packed, g, scale, dw, ws)
        self._h = c_void_p()
        call("ghnd_stem_wgrad_plan_create", ptr(x_packed), ptr(g), fmt_of(g.dtype), ptr(scale), ptr(dw), N, Hp,
             Wp, ptr(ws), nbytes, byref(self._h))
        self.desc = "stem_wgrad N%d %dx%d" % (N, Hp, Wp)
        self.flops = 2.0 * N * (Hp // 2) * (Wp // 2) * 64 * 147

    def run(self, stream=None):
        call("ghnd_stem_wgrad_plan_run", self._h, stream_ptr(stream))
        _count(2)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().ghnd_stem_wgrad_plan_destroy(h)
            except Exception:
                pass
            self._h = None


# ------------------------------------------------------------------------------------------------
# BatchNorm (training) and Adam
# ------------------------------------------------------------------------------------------------
def bn_stats(x, sums, planar=False, zeroed=False):
    """zeroed: the caller has zeroed `sums` (one memset per step over all such buffers, GHND_SUMS_ZEROED)."""
    flag = _lib.SUMS_ZEROED if zeroed else 0
    if planar:
        n, c, h, w = x.shape
        call("ghnd_bn_stats", ptr(x), 0, 1 | flag, n, h * w, c, ptr(sums), stream_ptr())
    else:
        n, h, w, c = x.shape
        call("ghnd_bn_stats", ptr(x), fmt_of(x.dtype), flag, n, h * w, c, ptr(sums), stream_ptr())
    _count()
    return sums


def bn_stats_finalize(x, sums, gamma, beta, eps, momentum, running_mean, running_var, nbt, scale_shift,
                      mean_invstd, planar=False, zeroed=False):
    """bn_stats + bn_finalize (count = the number of pixels of x) as one launch for planar tensors."""
    flag = _lib.SUMS_ZEROED if zeroed else 0
    if planar:
        n, c, h, w = x.shape
        fmt, pl = 0, 1 | flag
    else:
        n, h, w, c = x.shape
        fmt, pl = fmt_of(x.dtype), flag
    call("ghnd_bn_stats_finalize", ptr(x), fmt, pl, n, h * w, c, ptr(sums), ptr(gamma), ptr(beta), float(eps),
         float(momentum), ptr(running_mean), ptr(running_var), ptr(nbt), ptr(scale_shift), ptr(mean_invstd),
         stream_ptr())
    _count(1 if planar else 2)


def bn_finalize(sums, count, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, scale_shift,
                mean_invstd):
    call("ghnd_bn_finalize", ptr(sums), int(count), C, ptr(gamma), ptr(beta), float(eps), float(momentum),
         ptr(running_mean), ptr(running_var), ptr(nbt), ptr(scale_shift), ptr(mean_invstd), stream_ptr())
    _count()


def bn_eval_params(C, gamma, beta, running_mean, running_var, eps, scale_shift):
    call("ghnd_bn_eval_params", C, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), float(eps),
         ptr(scale_shift), stream_ptr())
    _count()
    return scale_shift


def bn_apply(x, y, scale_shift, relu, y2=None):
    n, h, w, c = x.shape
    call("ghnd_bn_apply", ptr(x), fmt_of(x.dtype), ptr(y), fmt_of(y.dtype), ptr(y2),
         fmt_of(y2.dtype) if y2 is not None else 0, n * h * w, c, ptr(scale_shift), int(bool(relu)),
         stream_ptr())
    _count()
    return y


def bn_finalize_apply(x, y, relu, sums, count, gamma, beta, eps, momentum, running_mean, running_var, nbt,
                      scale_shift, mean_invstd, y2=None):
    """bn_finalize + bn_apply in one launch (training forward); see ghnd_bn_finalize_apply."""
    n, h, w, c = x.shape
    call("ghnd_bn_finalize_apply", ptr(x), fmt_of(x.dtype), ptr(y), fmt_of(y.dtype), ptr(y2),
         fmt_of(y2.dtype) if y2 is not None else 0, n * h * w, c, int(bool(relu)), ptr(sums), int(count),
         ptr(gamma), ptr(beta), float(eps), float(momentum), ptr(running_mean), ptr(running_var), ptr(nbt),
         ptr(scale_shift), ptr(mean_invstd), stream_ptr())
    _count()
    return y


def convert16(x, y):
    call("ghnd_convert16", ptr(x), fmt_of(x.dtype), ptr(y), fmt_of(y.dtype), x.numel(), stream_ptr())
    _count()
    return y


def upsample_add(fine, coarse, out=None):
    """out = fine + nearest_upsample(coarse), NHWC 16-bit (the FPN top-down step)."""
    _need_cuda(fine, coarse)
    n, h, w, c = fine.shape
    _, hc, wc, _ = coarse.shape
    if out is None:
        out = torch.empty_like(fine)
    call("ghnd_upsample_add", ptr(fine), ptr(coarse), ptr(out), fmt_of(fine.dtype), n, h, w, hc, wc, c, stream_ptr())
    _count()
    return out


def adaptive_avgpool_nhwc16(x, oh, ow, out=None):
    """NHWC 16-bit [N,H,W,C] -> NHWC fp32 [N,oh,ow,C] (nn.AdaptiveAvgPool2d)."""
    _need_cuda(x)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=torch.float32, device=x.device)
    call("ghnd_adaptive_avgpool_nhwc16", ptr(x), fmt_of(x.dtype), n, h, w, c, ptr(out), oh, ow, stream_ptr())
    _count()
    return out


def small_conv_f32(x, w_rsck, scale, shift, relu, stride, out=None):
    """NHWC fp32 conv without padding + per-channel scale/shift (+ReLU); w_rsck [R,S,C,K] fp32."""
    _need_cuda(x, w_rsck)
    n, h, w, c = x.shape
    r, s, _, k = w_rsck.shape
    ho, wo = (h - r) // stride + 1, (w - s) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, k), dtype=torch.float32, device=x.device)
    call("ghnd_small_conv_f32", ptr(x), ptr(w_rsck), ptr(scale), ptr(shift), int(bool(relu)), ptr(out), n, h, w, c, k,
         r, s, stride, stream_ptr())
    _count()
    return out


def avgpool_linear(x, oh, ow, lw, lb, softmax, out=None):
    """NHWC fp32 [N,H,W,C] -> AdaptiveAvgPool2d(oh,ow) -> flatten (NCHW order) -> Linear -> [softmax]."""
    _need_cuda(x, lw)
    n, h, w, c = x.shape
    n_out = lw.shape[0]
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=x.device)
    call("ghnd_avgpool_linear", ptr(x), n, h, w, c, oh, ow, ptr(lw), ptr(lb), n_out, int(bool(softmax)), ptr(out),
         stream_ptr())
    _count()
    return out


def bn_bwd_reduce(dy, x, scale_shift, mean_invstd, relu, sums, planar=False, zeroed=False):
    flag = _lib.SUMS_ZEROED if zeroed else 0  # see bn_stats
    if planar:
        n, c, h, w = x.shape
        call("ghnd_bn_bwd_reduce", ptr(dy), 0, ptr(x), 0, 1 | flag, n, h * w, c, ptr(scale_shift),
             ptr(mean_invstd), int(bool(relu)), ptr(sums), stream_ptr())
    else:
        n, h, w, c = x.shape
        call("ghnd_bn_bwd_reduce", ptr(dy), fmt_of(dy.dtype), ptr(x), fmt_of(x.dtype), flag, n, h * w, c,
             ptr(scale_shift), ptr(mean_invstd), int(bool(relu)), ptr(sums), stream_ptr())
    _count()
    return sums


def bn_bwd_apply(dy, x, dx, gamma, scale_shift, mean_invstd, relu, sums, dgamma, dbeta, planar=False,
                 fused_sums=False):
    """fused_sums: `sums` = (sum g', sum g'*activation) written by the producing dgrad launch
    (ConvPlan(stats_mode=1)) instead of by bn_bwd_reduce."""
    if fused_sums:
        assert not planar
        n, h, w, c = x.shape
        call("ghnd_bn_bwd_apply_fused_sums", ptr(dy), fmt_of(dy.dtype), ptr(x), fmt_of(x.dtype), ptr(dx),
             fmt_of(dx.dtype), n, h * w, c, ptr(gamma), ptr(scale_shift), ptr(mean_invstd),
             int(bool(relu)), ptr(sums), ptr(dgamma), ptr(dbeta), stream_ptr())
        _count(1)
        return dx
    if planar:
        n, c, h, w = x.shape
        call("ghnd_bn_bwd_apply", ptr(dy), 0, ptr(x), 0, ptr(dx), 0, 1, n, h * w, c, ptr(gamma),
             ptr(scale_shift), ptr(mean_invstd), int(bool(relu)), ptr(sums), ptr(dgamma), ptr(dbeta),
             stream_ptr())
    else:
        n, h, w, c = x.shape
        call("ghnd_bn_bwd_apply", ptr(dy), fmt_of(dy.dtype), ptr(x), fmt_of(x.dtype), ptr(dx),
             fmt_of(dx.dtype), 0, n, h * w, c, ptr(gamma), ptr(scale_shift), ptr(mean_invstd),
             int(bool(relu)), ptr(sums), ptr(dgamma), ptr(dbeta), stream_ptr())
    _count(2 if planar else 1)
    return dx


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, grad_scale, step):
    call("ghnd_adam_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), float(lr),
         float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale), int(step),
         stream_ptr())
    _count()
