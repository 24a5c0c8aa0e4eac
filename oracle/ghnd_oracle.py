"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the reference's GHND hot path.

Nothing in the product package may import this module; it is used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the checker.

The reference (yoshitomo-matsubara/hnd-ghnd-object-detectors @7ebbe03) is pure Python whose
arithmetic lives in torch==1.3.1 / torchvision==0.4.2 library modules (Pipfile:7-8, not vendored).
This file restates that path with numpy (integer/byte work) and torch.nn.functional fp32 on CPU
(floating-point work), function by function, citing the reference lines each one follows.

Parity pin: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
UNMODIFIED reference from /root/reference (oracle/ref_loader.py) in the authoring container and
records its outputs on the deterministic weights of oracle/weights.py;
tests/test_oracle_golden.py checks this restatement against those vectors.  The reference itself
ships no tests or golden vectors for this path.
"""
import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# 8-bit quantizer -- src/myutils/pytorch/tensor_util.py:8-22
# --------------------------------------------------------------------------------------------


def quantize_tensor_np(x, num_bits=8, scale_mode="div"):
    """numpy fp32 restatement of quantize_tensor (tensor_util.py:8-18).

    scale_mode "div": scale = (max-min)/255 by IEEE division (torch CPU semantics, :12)
    scale_mode "recip": scale = (max-min)*fl32(1/255) (how torch evaluates tensor/python_scalar
    on a CUDA device).  Returns (q uint8, scale float32, zero_point int).
    """
    x = np.asarray(x, dtype=np.float32)
    f32 = np.float32
    qmin, qmax = f32(0.0), f32(2.0 ** num_bits - 1.0)
    mn, mx = x.min(), x.max()                                   # :11
    with np.errstate(all="ignore"):
        rng = f32(mx - mn)
        scale = f32(rng * f32(f32(1.0) / qmax)) if scale_mode == "recip" else f32(rng / qmax)  # :12
        izp = f32(qmin - f32(mn / scale))                       # :13
        if izp < qmin:                                          # :14
            zp = 0
        elif izp > qmax:
            zp = int(qmax)
        else:
            if np.isnan(izp):
                raise ValueError("cannot convert float NaN to integer")
            zp = int(izp)                                       # :15 truncation
        qx = f32(zp) + (x / scale).astype(np.float32)           # :16
        qx = np.rint(np.clip(qx, qmin, qmax))                   # :17 clamp, round half to even
    return qx.astype(np.uint8), scale, zp


def dequantize_tensor_np(q, scale, zero_point):
    """tensor_util.py:21-22: scale * (q.float() - zero_point)."""
    return (np.float32(scale) * (q.astype(np.float32) - np.float32(zero_point))).astype(np.float32)


# --------------------------------------------------------------------------------------------
# input transform -- src/models/org/rcnn.py:65-82 (normalize -> resize -> batch_images)
# --------------------------------------------------------------------------------------------
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


def resize_scale(h, w, size, max_size):
    """Scale factor of CustomRCNNTransform.resize (src/models/org/rcnn.py:29-42): min side -> `size`
    unless that pushes the max side over `max_size`."""
    mn, mx = float(min(h, w)), float(max(h, w))
    scale = size / mn
    if mx * scale > max_size:
        scale = max_size / mx
    return scale


def bilinear_resize_np(img, scale):
    """interpolate(img[None], scale_factor=scale, mode='bilinear', align_corners=False)[0]
    (rcnn.py:43-44) restated in numpy fp32 after ATen's upsample_bilinear2d: output size
    floor(double(in)*scale); source index rscale*(dst+0.5)-0.5 clamped at 0 with rscale =
    float(1/scale) (the scale factor is passed through, not recomputed from the sizes); second tap =
    first + (first < in-1); value = h0*(w0*a + w1*b) + h1*(w0*c + w1*d)."""
    import math
    x = np.asarray(img, dtype=np.float32)
    _, h, w = x.shape
    ho, wo = int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))
    rs = np.float32(1.0 / scale)

    def taps(n_out, n_in):
        src = rs * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5)
        src = np.maximum(src, np.float32(0)).astype(np.float32)
        i0 = np.minimum(src.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1).astype(np.float32), l1

    h0, h1, hl0, hl1 = taps(ho, h)
    w0, w1, wl0, wl1 = taps(wo, w)
    top = wl0[None, None, :] * x[:, h0][:, :, w0] + wl1[None, None, :] * x[:, h0][:, :, w1]
    bot = wl0[None, None, :] * x[:, h1][:, :, w0] + wl1[None, None, :] * x[:, h1][:, :, w1]
    return (hl0[None, :, None] * top + hl1[None, :, None] * bot).astype(np.float32)


def transform_batch(images, size_divisible=32, sizes=None, max_size=1333):
    """normalize -> resize -> zero-pad to a common size rounded up to 32 (rcnn.py:65-82;
    torchvision GeneralizedRCNNTransform.normalize / batch_images).  `sizes` = per-image target min
    side (fixed_sizes of DistillationBox, tool.py:44-49, or min_size[-1] in eval); None = the images
    are already at network scale (resize is the identity: interpolate with scale_factor 1)."""
    dev = images[0].device  # the bench's "stock torch on the same B200" leg runs this port on cuda
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32, device=dev)[:, None, None]
    std = torch.tensor(IMAGE_STD, dtype=torch.float32, device=dev)[:, None, None]
    imgs = [(im - mean) / std for im in images]
    if sizes is not None:
        scaled = []
        for im, size in zip(imgs, sizes):
            sc = resize_scale(im.shape[1], im.shape[2], size, max_size)
            scaled.append(im if sc == 1.0 else torch.from_numpy(bilinear_resize_np(im.cpu().numpy(), sc)).to(dev))
        imgs = scaled
    hmax = max(im.shape[1] for im in imgs)
    wmax = max(im.shape[2] for im in imgs)
    hp = (hmax + size_divisible - 1) // size_divisible * size_divisible
    wp = (wmax + size_divisible - 1) // size_divisible * size_divisible
    out = torch.zeros(len(imgs), 3, hp, wp, dtype=torch.float32, device=dev)
    for i, im in enumerate(imgs):
        out[i, :, :im.shape[1], :im.shape[2]] = im
    return out


# --------------------------------------------------------------------------------------------
# backbone pieces
# --------------------------------------------------------------------------------------------
FROZEN_BN_EPS = 1e-5  # torchvision 0.26 FrozenBatchNorm2d default (0.4.2 had no eps)


def frozen_bn(x, sd, prefix):
    """torchvision.ops.misc.FrozenBatchNorm2d.forward."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    scale = w * (rv + FROZEN_BN_EPS).rsqrt()
    bias = b - rm * scale
    return x * scale[None, :, None, None] + bias[None, :, None, None]


def stem_forward(x, sd, prefix="backbone.body."):
    """conv1 7x7 s2 p3 -> FrozenBN -> ReLU -> MaxPool 3x3 s2 p1 (custom/resnet.py:26-30,96-99)."""
    x = F.conv2d(x, sd[prefix + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(frozen_bn(x, sd, prefix + "bn1"))
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def bottleneck_forward(x, sd, prefix, stride):
    """torchvision.models.resnet.Bottleneck.forward with FrozenBN (stride on the 3x3)."""
    identity = x
    out = F.relu(frozen_bn(F.conv2d(x, sd[prefix + ".conv1.weight"]), sd, prefix + ".bn1"))
    out = F.relu(frozen_bn(F.conv2d(out, sd[prefix + ".conv2.weight"], None, stride, 1), sd, prefix + ".bn2"))
    out = frozen_bn(F.conv2d(out, sd[prefix + ".conv3.weight"]), sd, prefix + ".bn3")
    if (prefix + ".downsample.0.weight") in sd:
        identity = frozen_bn(F.conv2d(x, sd[prefix + ".downsample.0.weight"], None, stride), sd,
                             prefix + ".downsample.1")
    return F.relu(out + identity)


RESNET50_BLOCKS = {"layer1": 3, "layer2": 4, "layer3": 6, "layer4": 3}


def frozen_layer_forward(x, sd, name, prefix="backbone.body."):
    for b in range(RESNET50_BLOCKS[name]):
        stride = 2 if (b == 0 and name != "layer1") else 1
        x = bottleneck_forward(x, sd, "%s%s.%d" % (prefix, name, b), stride)
    return x


def _bn(x, sd, prefix, training, eps=1e-5, momentum=0.1, update=None):
    """nn.BatchNorm2d.forward; `update` (a dict) receives the new running stats in training."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, eps)
        if update is not None:
            update[prefix + ".running_mean"] = rm2
            update[prefix + ".running_var"] = rv2
        return y
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], False, momentum, eps)


def student_encoder_forward(x, sd, prefix, training, update=None):
    """Bottleneck4LargeResNet.encoder (src/models/mimic/resnet_layer.py:42-51)."""
    e = prefix + ".encoder.encoder."
    x = _bn(F.conv2d(x, sd[e + "0.weight"], None, 1, 1), sd, e + "1", training, update=update)
    x = F.relu(_bn(F.conv2d(x, sd[e + "2.weight"], None, 1, 1), sd, e + "3", training, update=update))
    x = _bn(F.conv2d(x, sd[e + "5.weight"], None, 1, 1), sd, e + "6", training, update=update)
    return F.conv2d(x, sd[e + "7.weight"], None, 1, 1)


def student_decoder_forward(z, sd, prefix, training, update=None):
    """Bottleneck4LargeResNet.decoder (resnet_layer.py:52-65)."""
    d = prefix + ".decoder."
    z = F.relu(_bn(z, sd, d + "0", training, update=update))
    z = _bn(F.conv2d(z, sd[d + "2.weight"]), sd, d + "3", training, update=update)
    z = F.relu(_bn(F.conv2d(z, sd[d + "4.weight"]), sd, d + "5", training, update=update))
    z = _bn(F.conv2d(z, sd[d + "7.weight"]), sd, d + "8", training, update=update)
    return F.relu(_bn(F.conv2d(z, sd[d + "9.weight"]), sd, d + "10", training, update=update))


def student_layer1_forward(x, sd, prefix="backbone.body.layer1", training=True, quantize_bits=None,
                           update=None):
    """BottleneckBase4Ext.forward (src/models/mimic/base.py:50-58): encoder -> [eval-only
    quantize/dequantize] -> decoder."""
    z = student_encoder_forward(x, sd, prefix, training, update)
    if quantize_bits is not None and not training:
        q, scale, zp = quantize_tensor_np(z.detach().numpy(), quantize_bits)
        z = torch.from_numpy(dequantize_tensor_np(q, scale, zp))
    return student_decoder_forward(z, sd, prefix, training, update)


def backbone_features(x, sd, student, training=False, update=None, prefix="backbone.body."):
    """IntermediateLayerGetter over conv1..layer4 (src/models/org/rcnn.py:399-414); returns the
    layer1..layer4 outputs the DistillationBox hooks observe (src/distillation/tool.py:19-35)."""
    feats = {}
    x = stem_forward(x, sd, prefix)
    if student:
        x = student_layer1_forward(x, sd, prefix + "layer1", training, update=update)
    else:
        x = frozen_layer_forward(x, sd, "layer1", prefix)
    feats["layer1"] = x
    for name in ("layer2", "layer3", "layer4"):
        x = frozen_layer_forward(x, sd, name, prefix)
        feats[name] = x
    return feats


# --------------------------------------------------------------------------------------------
# loss + step -- src/distillation/loss.py:25-34, src/distillation/tool.py:40-61
# --------------------------------------------------------------------------------------------


def ghnd_loss(teacher_feats, student_feats, levels=("layer1", "layer2", "layer3", "layer4"),
              factors=None):
    """sum_l factor_l * MSELoss(reduction='sum')(teacher_l, student_l)."""
    total = 0.0
    per_level = {}
    for lv in levels:
        f = 1.0 if factors is None else factors[lv]
        term = F.mse_loss(teacher_feats[lv], student_feats[lv], reduction="sum") * f
        per_level[lv] = term
        total = total + term
    return total, per_level


TRAINABLE_SUFFIXES = ("conv1.weight",)


def trainable_names(student_sd, prefix="backbone.body."):
    """Exactly what mimic_runner.freeze_modules leaves trainable."""
    names = [prefix + "conv1.weight"]
    for k in student_sd:
        if k.startswith(prefix + "layer1.") and (k.endswith(".weight") or k.endswith(".bias")):
            names.append(k)
    return names


def distill_step(teacher_sd, student_sd, images, levels=("layer1", "layer2", "layer3", "layer4"),
                 sizes=None, max_size=1333):
    """One DistillationBox.forward + backward (tool.py:40-61, mimic_runner.py:51-53).
    Returns loss, per-level terms, teacher/student features, grads of the trainable tensors and the
    updated BN running stats.  `sizes` = the per-image fixed_sizes of the Keypoint path
    (tool.py:44-49); both models see the same resized batch."""
    x = transform_batch(images, sizes=sizes, max_size=max_size)
    with torch.no_grad():
        t_feats = backbone_features(x, teacher_sd, student=False)
    names = trainable_names(student_sd)
    sd = dict(student_sd)
    leaves = {}
    for n in names:
        leaves[n] = student_sd[n].detach().clone().requires_grad_(True)
        sd[n] = leaves[n]
    update = {}
    s_feats = backbone_features(x, sd, student=True, training=True, update=update)
    loss, per_level = ghnd_loss(t_feats, s_feats, levels)
    grads = torch.autograd.grad(loss, [leaves[n] for n in names])
    return {
        "loss": loss.detach(),
        "per_level": {k: v.detach() for k, v in per_level.items()},
        "teacher": t_feats,
        "student": {k: v.detach() for k, v in s_feats.items()},
        "grads": dict(zip(names, grads)),
        "bn_update": update,
    }


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (mimic_runner.py:54 via func_util.get_optimizer)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    return p - (lr / bc1) * (m / denom), m, v


def encode_head(images, student_sd, num_bits=8):
    """RcnnHead.forward (src/models/mimic/split_rcnn.py:23-37): transform -> stem -> encoder (eval)
    -> Quantizer.  Returns (q, scale, zp, z, padded batch shape)."""
    x = transform_batch(images)
    with torch.no_grad():
        z = student_encoder_forward(stem_forward(x, student_sd), student_sd, "backbone.body.layer1",
                                    training=False)
    q, scale, zp = quantize_tensor_np(z.cpu().numpy(), num_bits)
    return q, scale, zp, z, tuple(x.shape)
