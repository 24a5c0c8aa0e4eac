"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights shared by the golden generator,
the oracle tests and the GPU parity tests (no checkpoints exist offline).

Every tensor is drawn from its own torch.Generator seeded by a stable hash of its name, so a test
can rebuild exactly the weights the golden fixtures were recorded with, without the reference.
Statistics follow the reference's init (Kaiming-normal fan_out convs, custom/resnet.py:55-60) but
the (Frozen)BatchNorm tensors are made non-trivial so that folding bugs cannot hide.
"""
import zlib

import torch

RESNET50_BLOCKS = {"layer1": 3, "layer2": 4, "layer3": 6, "layer4": 3}
PLANES = {"layer1": 64, "layer2": 128, "layer3": 256, "layer4": 512}


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _conv(name, o, i, k, seed, gain=1.0):
    std = gain * (2.0 / (o * k * k)) ** 0.5
    return torch.randn(o, i, k, k, generator=_gen(name, seed)) * std


def _norm(sd, prefix, c, seed, frozen):
    g = _gen(prefix, seed)
    sd[prefix + ".weight"] = 0.75 + 0.5 * torch.rand(c, generator=g)
    sd[prefix + ".bias"] = 0.2 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    if not frozen:
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _bottleneck(sd, prefix, inplanes, planes, downsample, seed):
    sd[prefix + ".conv1.weight"] = _conv(prefix + ".conv1", planes, inplanes, 1, seed)
    _norm(sd, prefix + ".bn1", planes, seed, True)
    sd[prefix + ".conv2.weight"] = _conv(prefix + ".conv2", planes, planes, 3, seed)
    _norm(sd, prefix + ".bn2", planes, seed, True)
    sd[prefix + ".conv3.weight"] = _conv(prefix + ".conv3", planes * 4, planes, 1, seed, gain=0.5)
    _norm(sd, prefix + ".bn3", planes * 4, seed, True)
    if downsample:
        sd[prefix + ".downsample.0.weight"] = _conv(prefix + ".downsample.0", planes * 4, inplanes, 1,
                                                    seed, gain=0.7)
        _norm(sd, prefix + ".downsample.1", planes * 4, seed, True)


def frozen_body(seed=0, prefix="backbone.body.", with_layer1=True, stem_name="teacher"):
    """conv1/bn1 + torchvision-ResNet-50 layer1..4 tensors (names as in the reference state_dict)."""
    sd = {}
    sd[prefix + "conv1.weight"] = _conv(stem_name + ".conv1", 64, 3, 7, seed)
    _norm(sd, prefix + "bn1", 64, seed, True)
    inplanes = 64
    for name in ("layer1", "layer2", "layer3", "layer4"):
        planes = PLANES[name]
        for b in range(RESNET50_BLOCKS[name]):
            if name != "layer1" or with_layer1:
                _bottleneck(sd, "%s%s.%d" % (prefix, name, b), inplanes, planes, b == 0, seed)
            inplanes = planes * 4
    return sd


def student_layer1(bch=3, seed=0, prefix="backbone.body.layer1."):
    """Bottleneck4LargeResNet tensors (src/models/mimic/resnet_layer.py:42-65)."""
    sd = {}
    e, d = prefix + "encoder.encoder.", prefix + "decoder."
    enc = [("0", 64, 64), ("2", 256, 64), ("5", 64, 256), ("7", bch, 64)]
    for idx, o, i in enc:
        sd[e + idx + ".weight"] = _conv(e + idx, o, i, 2, seed)
    for idx, c in (("1", 64), ("3", 256), ("6", 64)):
        _norm(sd, e + idx, c, seed, False)
    _norm(sd, d + "0", bch, seed, False)
    dec = [("2", 64, bch), ("4", 128, 64), ("7", 256, 128), ("9", 256, 256)]
    for idx, o, i in dec:
        sd[d + idx + ".weight"] = _conv(d + idx, o, i, 2, seed)
    for idx, c in (("3", 64), ("5", 128), ("8", 256), ("10", 256)):
        _norm(sd, d + idx, c, seed, False)
    return sd


def teacher_student(bch=3, seed=0):
    """(teacher_sd, student_sd): layer2-4 and bn1 shared, the student's conv1 differs (it is
    trainable in the reference)."""
    teacher = frozen_body(seed)
    student = {k: v.clone() for k, v in frozen_body(seed, with_layer1=False, stem_name="student").items()}
    student.update(student_layer1(bch, seed))
    return teacher, student
