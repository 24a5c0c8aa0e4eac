"""TEST INFRASTRUCTURE ONLY -- storage-precision emulation of the GHND step on top of the fp32 oracle.

`oracle/ghnd_oracle.py` restates the reference in fp32.  The CUDA path computes the same graph but
STORES activations and weights as fp16 and gradients as bf16; every
convolution accumulates in fp32.  Against the fp32 oracle the end-to-end gradients therefore differ
by 5-10 % relative L2 -- almost all of it ReLU-mask flips of activations within 16-bit rounding of
zero -- which is too loose to tell "rounding moved a mask" from "a backward kernel is wrong".

This module keeps the oracle's graph and torch.autograd's chain rule and only injects rounding at
exactly the tensors the engine stores (engine.py / conv_tc.cu epilogues):

  * Q(x)   forward: x -> fp16 -> fp32         backward: identity        (a stored activation)
  * G(x)   forward: identity                  backward: g -> bf16 -> fp32 (a stored gradient)
  * conv   forward with fp16(w_eff), data gradient with bf16(w_eff), weight gradient from the bf16
           copy of the input activation (the tcgen05 dW operands), all with fp32 accumulation;
           w_eff = w * FrozenBN scale for the frozen convs (ghnd_pack_weight)
  * frozen-conv epilogue: h = fp16(acc); h = fp16(h + fp16(bias)); h = fp16(h + residual); ReLU
    (epi_half_rows: packed half2 arithmetic after one conversion of the accumulator)
  * student layer1: raw = fp16(acc); BatchNorm (batch statistics of the stored raw) in fp32;
    out = fp16(.), its bf16 copy feeds the next unit's dW; the bottleneck z and the narrow convs
    around it stay fp32 (planar fp32 tensors, fp32-split mma.sync operands)
  * loss gradient = bf16(2 * factor * (s - t)), masked by s > 0 on the top level (sse.cu)

Free-running, the emulation agrees with the engine to ~1e-5 per kernel, but 16-bit rounding is chaotic
along a deep chain: a 1-ulp difference in 0.2 % of one layer's outputs (accumulation order) re-rounds a few
per cent of the next layer's, and after ~6 convolutions two CORRECT implementations sit at the fp16
rounding floor from each other (relative L2 ~1e-3, measured at 800x1333 with scripts/debug/emu_vs_engine.py).
Their ReLU masks then differ in the same ~0.2 % of positions as against the fp32 oracle, so end-to-end
gradients differ by percents no matter how exact the backward kernels are.  To test the BACKWARD at step
level the emulation can therefore be teacher-forced: `force` maps storage points to the tensors the engine
actually stored in its forward pass; the emulated forward takes those VALUES (hence the engine's masks,
arg-maxima and BatchNorm statistics) while autograd still differentiates the oracle's graph.  The GPU tests
gate gradients at <= 1e-2 relative L2 against the forced emulation and REPORT the distances to the
free-running emulation and to the fp32 oracle next to it.
Reference lines restated: the same as ghnd_oracle.py (tool.py:40-61, resnet_layer.py:42-65,
custom/resnet.py:26-30,96-99, loss.py:25-34)."""
import torch
import torch.nn.functional as F

from . import ghnd_oracle as O


def r16(x):
    return x.half().float()


def rbf(x):
    return x.bfloat16().float()


class _Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return r16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _G(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return rbf(g)


Q, G = _Q.apply, _G.apply

# teacher forcing (see the module docstring): storage-point name -> tensor stored by the engine (NCHW fp32)
_FORCE = {}
_GRADS = None      # set to a dict to capture the gradient arriving at every named storage point
_FORCE_GRAD = {}   # storage-point name -> gradient tensor that REPLACES the emulated one there


def _capture(t, key):
    if key and t.requires_grad:
        if _GRADS is not None:
            t.register_hook(lambda g, k=key: _GRADS.__setitem__(k, g.detach().clone()))
        if key in _FORCE_GRAD:
            t.register_hook(lambda g, k=key: _FORCE_GRAD[k].to(g.dtype))
    return t


def _forced(x, key):
    """Value of the engine's stored tensor, gradient of the emulated one."""
    f = _FORCE.get(key)
    _capture(x, key)
    if f is None:
        return x
    assert f.shape == x.shape, (key, tuple(f.shape), tuple(x.shape))
    return f + (x - x.detach())


def _forced_relu(pre, key):
    """ReLU whose output value AND mask are the engine's stored post-ReLU tensor."""
    f = _FORCE.get(key)
    _capture(pre, key)
    if f is None:
        return F.relu(pre)
    assert f.shape == pre.shape, (key, tuple(f.shape), tuple(pre.shape))
    m = (f > 0).to(pre.dtype)
    return f + (pre * m - (pre * m).detach())


class _Conv(torch.autograd.Function):
    """y = conv(x, fp16(w)); dx = conv_T(g, bf16(w)); dw = corr(xg, g) with xg = the bf16 copy of x."""

    @staticmethod
    def forward(ctx, x, xg, w, stride, pad):
        ctx.save_for_backward(xg if xg is not None else x, w)
        ctx.x_shape, ctx.stride, ctx.pad = x.shape, stride, pad
        return F.conv2d(x, r16(w), None, stride, pad)

    @staticmethod
    def backward(ctx, g):
        xg, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.nn.grad.conv2d_input(ctx.x_shape, rbf(w), g, ctx.stride, ctx.pad)
        if ctx.needs_input_grad[2]:
            gw = torch.nn.grad.conv2d_weight(xg, w.shape, g, ctx.stride, ctx.pad)
        return gx, None, gw, None, None


def conv16(x, w, stride=1, pad=0, xg=None):
    return _Conv.apply(x, xg, w, stride, pad)


def _frozen_scale_shift(sd, prefix):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    scale = w * (rv + O.FROZEN_BN_EPS).rsqrt()
    return scale, b - rm * scale


def frozen_conv(x, sd, conv, bn, stride=1, pad=0, relu=True, residual=None, xg=None, name=None):
    """conv + folded FrozenBN through the packed-half2 epilogue (conv_tc.cu epi_half_rows).
    `name`: storage point of the result for teacher forcing."""
    scale, shift = _frozen_scale_shift(sd, bn)
    h = Q(conv16(x, sd[conv + ".weight"] * scale[:, None, None, None], stride, pad, xg))
    h = Q(h + r16(shift)[None, :, None, None])
    if residual is not None:
        h = Q(h + residual)
    return _forced_relu(h, name) if relu else _forced(h, name)


def stem16(x16, sd, prefix="backbone.body."):
    """packed fp16 image -> conv1 (+FrozenBN, ReLU) -> max-pool; conv1's dW uses the bf16 image copy."""
    c = frozen_conv(x16, sd, prefix + "conv1", prefix + "bn1", 2, 3, True, xg=rbf(x16), name=prefix + "conv1")
    return F.max_pool2d(G(c), kernel_size=3, stride=2, padding=1)


def bottleneck16(x, sd, prefix, stride):
    """torchvision Bottleneck on the engine's storage points (engine.BottleneckRunner).  Returns
    (out, tap): `tap` is the node of x on which a loss term attached to x must hang so that its
    gradient enters the same bf16 sum as in the dgrad epilogue (residual operand of conv1's dgrad)."""
    ds = (prefix + ".downsample.0.weight") in sd
    x0 = G(x)                       # the stored gradient w.r.t. x (after the last accumulating launch)
    xi = G(x0) if ds else x0        # downsample blocks: inner sum rbf(rbf(acc_conv1) + loss_grad) first
    a1 = frozen_conv(G(xi), sd, prefix + ".conv1", prefix + ".bn1", name=prefix + ".a1")
    a2 = frozen_conv(G(a1), sd, prefix + ".conv2", prefix + ".bn2", stride, 1, name=prefix + ".a2")
    if ds:
        idn = frozen_conv(G(x0), sd, prefix + ".downsample.0", prefix + ".downsample.1", stride, 0, relu=False,
                          name=prefix + ".idn")
    else:
        idn = xi
    out = frozen_conv(G(a2), sd, prefix + ".conv3", prefix + ".bn3", relu=True, residual=idn, name=prefix + ".out")
    return out, xi


def frozen_layer16(x, sd, name, prefix="backbone.body."):
    tap = None
    for b in range(O.RESNET50_BLOCKS[name]):
        stride = 2 if (b == 0 and name != "layer1") else 1
        x, t = bottleneck16(x, sd, "%s%s.%d" % (prefix, name, b), stride)
        if b == 0:
            tap = t
    return x, tap


def _bn_train(x, sd, prefix, eps=1e-5):
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, eps)


def _bn_eval(x, sd, prefix, eps=1e-5):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, eps)


def _wide_unit(x, xg, sd, conv, bn, pad, relu, training=True):
    """engine._WideUnit: conv (statistics in the epilogue) -> BatchNorm(batch stats) [-> ReLU].
    Returns (stored fp16 output, its bf16 copy for the next unit's dW)."""
    if not training:  # eval: BN folded into the conv (scale into fp16 weights, shift as fp16 bias)
        rv, rm = sd[bn + ".running_var"], sd[bn + ".running_mean"]
        sc = sd[bn + ".weight"] / torch.sqrt(rv + 1e-5)
        h = Q(conv16(x, sd[conv + ".weight"] * sc[:, None, None, None], 1, pad))
        h = Q(h + r16(sd[bn + ".bias"] - rm * sc)[None, :, None, None])
        return (F.relu(h) if relu else h), None
    raw = G(_forced(Q(conv16(G(x), sd[conv + ".weight"], 1, pad, xg)), conv + ".raw"))
    a = _bn_train(raw, sd, bn)
    if relu:
        a = F.relu(a)
    return _forced(Q(a), conv + ".out"), rbf(a.detach())


def student_layer1_16(x, sd, prefix="backbone.body.layer1", training=True):
    """Bottleneck4LargeResNet on the engine's storage points (engine.StudentLayer1Runner)."""
    e, d = prefix + ".encoder.encoder.", prefix + ".decoder."
    y, yg = _wide_unit(x, rbf(x.detach()), sd, e + "0", e + "1", 1, False, training)
    y, yg = _wide_unit(y, yg, sd, e + "2", e + "3", 1, True, training)
    y, _ = _wide_unit(y, yg, sd, e + "5", e + "6", 1, False, training)
    z = F.conv2d(G(y) if training else y, sd[e + "7.weight"], None, 1, 1)      # fp32 planar bottleneck
    bn = _bn_train if training else _bn_eval
    a = F.relu(bn(z, sd, d + "0"))
    raw3 = _forced(Q(F.conv2d(a, sd[d + "2.weight"])), d + "2.raw")
    if training:
        raw3 = G(raw3)
    a3 = bn(raw3, sd, d + "3")
    y, yg = _forced(Q(a3), d + "2.out"), rbf(a3.detach())
    y, yg = _wide_unit(y, yg, sd, d + "4", d + "5", 0, True, training)
    y, yg = _wide_unit(y, yg, sd, d + "7", d + "8", 0, False, training)
    y, _ = _wide_unit(y, yg, sd, d + "9", d + "10", 0, True, training)
    return y


def backbone_features16(x16, sd, student, training=False, prefix="backbone.body."):
    """-> (features per level, tap per level): the loss term of a level hangs on its tap."""
    feats, taps = {}, {}
    x = stem16(x16, sd, prefix)
    if student:
        x = student_layer1_16(x, sd, prefix + "layer1", training)
    else:
        x, _ = frozen_layer16(x, sd, "layer1", prefix)
    feats["layer1"] = x
    prev = "layer1"
    for name in ("layer2", "layer3", "layer4"):
        x, tap = frozen_layer16(x, sd, name, prefix)
        taps[prev] = tap
        feats[name] = x
        prev = name
    return feats, taps


def distill_step16(teacher_sd, student_sd, images, levels=("layer1", "layer2", "layer3", "layer4"),
                   sizes=None, max_size=1333, force=None, teacher_feats=None, force_grad=None, capture=None):
    """One GHND step with the engine's storage precision; same return layout as O.distill_step.
    force / teacher_feats: teacher forcing with the tensors the engine stored (module docstring);
    keys of `force` are the storage-point names used above (e.g. 'backbone.body.layer2.0.a1',
    'backbone.body.layer1.decoder.9.raw', 'backbone.body.conv1').  force_grad: the same for stored
    GRADIENTS (the backward continues from the engine's tensor at that point); capture: a dict that
    receives the emulated gradient at every named storage point."""
    global _GRADS
    _FORCE.clear()
    _FORCE_GRAD.clear()
    if force:
        _FORCE.update(force)
    if force_grad:
        _FORCE_GRAD.update(force_grad)
    _GRADS = capture
    try:
        return _distill_step16(teacher_sd, student_sd, images, levels, sizes, max_size, teacher_feats)
    finally:
        _FORCE.clear()
        _FORCE_GRAD.clear()
        _GRADS = None


def _distill_step16(teacher_sd, student_sd, images, levels, sizes, max_size, teacher_feats):
    x16 = r16(O.transform_batch(images, sizes=sizes, max_size=max_size))
    if teacher_feats is not None:
        t_feats = teacher_feats
    else:
        saved = dict(_FORCE)
        _FORCE.clear()  # the storage-point names are the student's
        with torch.no_grad():
            t_feats, _ = backbone_features16(x16, teacher_sd, student=False)
        _FORCE.update(saved)
    names = O.trainable_names(student_sd)
    sd = dict(student_sd)
    leaves = {}
    for n in names:
        leaves[n] = student_sd[n].detach().clone().requires_grad_(True)
        sd[n] = leaves[n]
    s_feats, taps = backbone_features16(x16, sd, student=True, training=True)
    order = ("layer1", "layer2", "layer3", "layer4")
    top = max(order.index(l) for l in levels)
    total, per_level = 0.0, {}
    for lv in levels:
        if order.index(lv) == top:
            s = G(s_feats[lv])           # loss gradient stored as bf16, then the ReLU mask
        else:
            s = G(taps[lv])              # residual operand of the next layer's first dgrad
        term = ((t_feats[lv].double() - s.double()) ** 2).sum()
        per_level[lv] = term
        total = total + term
    grads = torch.autograd.grad(total, [leaves[n] for n in names])
    return {
        "loss": total.detach().float(),
        "per_level": {k: v.detach().float() for k, v in per_level.items()},
        "teacher": t_feats,
        "student": {k: v.detach() for k, v in s_feats.items()},
        "grads": dict(zip(names, [g.float() for g in grads])),
    }
