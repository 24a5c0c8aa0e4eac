"""Student head modules -- drop-in mirror of the reference's src/models/mimic/resnet_layer.py and
src/models/mimic/base.py.

Same class names, constructor arguments, attributes (.encoder/.decoder/.bottleneck_transformer/
.use_bottleneck_transformer) and state_dict keys (encoder.encoder.{0..7}.*, decoder.{0..10}.*) so
released checkpoints load with strict=True.  The nn.Conv2d / nn.BatchNorm2d children only HOLD the
parameters; forward() runs the hand-written sm_100a kernels through StudentLayer1Runner (engine.py)
and raises if the input is not on a CUDA device (no CPU fallback).
"""
import torch
from torch import nn

from . import _lib, ops


def apply_bottleneck_transformer(layer1, z):
    """base.py:55-57: z, _ = bottleneck_transformer(z, target=None); z = z.to(device).  When
    `layer1.data_logging` is set the first wire object of the chain (e.g. the QuantizedTensor between
    Quantizer and Dequantizer) is kept in `layer1.last_bottleneck` for byte accounting."""
    device = z.device
    tr = layer1.bottleneck_transformer
    if getattr(layer1, 'data_logging', False) and hasattr(tr, 'transforms'):
        layer1.last_bottleneck = None
        for t in tr.transforms:
            z, _ = t(z, None)
            if layer1.last_bottleneck is None and not isinstance(z, torch.Tensor):
                layer1.last_bottleneck = z
        if layer1.last_bottleneck is None:
            layer1.last_bottleneck = z
    else:
        z, _ = tr(z, target=None)
    return z.to(device)


class ExtEncoder(nn.Module):
    """base.py:7-26: the encoder stack plus the optional neural filter `ext_classifier`
    (models/ext/classifier.py) that decides from the stem output whether an image is worth encoding:
    in eval mode with batch 1, `ext_z[0][1] < threshold` stops inference before the encoder runs."""

    def __init__(self, encoder, ext_classifier=None, ext_config=None):
        super().__init__()
        self.encoder = encoder
        self.ext_classifier = ext_classifier
        self.threshold = ext_config['threshold'] if ext_config is not None else None

    def forward(self, x):
        raise _lib.GhndError("ExtEncoder is executed by its parent Bottleneck4LargeResNet")

    def filter_decision(self, x_nhwc16):
        """forward_with_ext's test (base.py:13-16) on the NHWC 16-bit stem output: returns (skip, ext_z);
        skip is True when the filter says nothing of interest is in the (single) image."""
        ext_z = self.ext_classifier.forward_nhwc16(x_nhwc16)
        skip = (not self.training) and ext_z.shape[0] == 1 and bool(ext_z[0][1] < self.threshold)
        return skip, ext_z

    def get_ext_classifier(self):
        return self.ext_classifier


class _Layer1Function(torch.autograd.Function):
    """Training-mode forward/backward of the whole bottleneck layer through the CUDA runner."""

    @staticmethod
    def forward(ctx, module, x, *params):
        runner = module._runner(x.shape, train=True)
        ops.to_nhwc16_into(x, runner.x)
        out = runner.forward()
        ctx.module, ctx.runner = module, runner
        return ops.to_nchw_f32(out)

    @staticmethod
    def backward(ctx, g):
        module, runner = ctx.module, ctx.runner
        grads = module._grad_views(runner)
        ops.to_nhwc16_into(g.contiguous(), runner.g_in)
        runner.backward()
        gx = ops.to_nchw_f32(runner.g_x)
        names = [n for n, _ in module.named_parameters()]
        return (None, gx) + tuple(grads["layer1." + n].clone() for n in names)


class BottleneckBase4Ext(nn.Module):
    """base.py:28-61: encoder -> [eval & use_bottleneck_transformer: transformer(z)] -> decoder."""

    def __init__(self, encoder, decoder, bottleneck_transformer=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.bottleneck_transformer = bottleneck_transformer
        from .transformer import DataLogger
        self.data_logging = isinstance(bottleneck_transformer, DataLogger)
        self.last_bottleneck = None
        self.uses_ext_encoder = isinstance(encoder, ExtEncoder) and encoder.ext_classifier is not None
        self.use_bottleneck_transformer = False
        self._runners = {}
        self.act_dtype = torch.float16
        self.grad_dtype = torch.bfloat16

    # ---- plumbing -------------------------------------------------------------------------
    def _runner(self, shape, train):
        from .engine import StudentLayer1Runner, _empty
        n, c, h, w = shape
        key = (n, h, w, bool(train))
        r = self._runners.get(key)
        if r is None:
            dev = self.decoder[0].weight.device
            x = _empty((n, h, w, 64), self.act_dtype, dev)
            r = StudentLayer1Runner(self, x, n, h, w, self.act_dtype, self.grad_dtype, train)
            if train:
                r.g_in = _empty((n, h, w, 256), self.grad_dtype, dev)
                views = {}
                for name, p in self.named_parameters():
                    views["layer1." + name] = torch.zeros_like(p, dtype=torch.float32)
                r.plan_backward(r.g_in, views, "layer1.")
                r.grad_views = views
            self._runners[key] = r
        return r

    def _grad_views(self, runner):
        return runner.grad_views

    def _check(self, x):
        if not x.is_cuda:
            raise _lib.GhndError("Bottleneck4LargeResNet runs on CUDA only (sm_100a kernels, no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 64:
            raise ValueError("expected input of shape [N, 64, H, W], got %s" % (tuple(x.shape),))

    def forward(self, x):
        self._check(x)
        if self.uses_ext_encoder and self.training:
            raise _lib.GhndError("training with the neural filter attached (ext_runner.py) is outside the "
                                 "B200 hot path")
        if self.training and torch.is_grad_enabled():
            return _Layer1Function.apply(self, x, *list(self.parameters()))
        runner = self._runner(x.shape, train=self.training)
        ops.to_nhwc16_into(x, runner.x)
        ext_z = None
        if self.uses_ext_encoder:  # base.py:38-48 forward_ext: returns (decoder output or None, ext_z)
            skip, ext_z = self.encoder.filter_decision(runner.x)
            if skip:
                if self.data_logging:
                    self.bottleneck_transformer(None, target=None)
                return None, ext_z
        z = runner.forward_encoder()
        if not self.training and self.bottleneck_transformer is not None and self.use_bottleneck_transformer:
            z = apply_bottleneck_transformer(self, z)  # base.py:55-57
        out = ops.to_nchw_f32(runner.forward_decoder(z.contiguous()))
        return (out, ext_z) if self.uses_ext_encoder else out

    def get_ext_classifier(self):
        return self.encoder.get_ext_classifier() if isinstance(self.encoder, ExtEncoder) else None


def _make_encoder_decoder(bottleneck_channel):
    encoder = nn.Sequential(
        nn.Conv2d(64, 64, kernel_size=2, padding=1, bias=False),
        nn.BatchNorm2d(64),
        nn.Conv2d(64, 256, kernel_size=2, padding=1, bias=False),
        nn.BatchNorm2d(256),
        nn.ReLU(inplace=True),
        nn.Conv2d(256, 64, kernel_size=2, padding=1, bias=False),
        nn.BatchNorm2d(64),
        nn.Conv2d(64, bottleneck_channel, kernel_size=2, padding=1, bias=False)
    )
    decoder = nn.Sequential(
        nn.BatchNorm2d(bottleneck_channel),
        nn.ReLU(inplace=True),
        nn.Conv2d(bottleneck_channel, 64, kernel_size=2, bias=False),
        nn.BatchNorm2d(64),
        nn.Conv2d(64, 128, kernel_size=2, bias=False),
        nn.BatchNorm2d(128),
        nn.ReLU(inplace=True),
        nn.Conv2d(128, 256, kernel_size=2, bias=False),
        nn.BatchNorm2d(256),
        nn.Conv2d(256, 256, kernel_size=2, bias=False),
        nn.BatchNorm2d(256),
        nn.ReLU(inplace=True)
    )
    return encoder, decoder


class Bottleneck4LargeResNet(BottleneckBase4Ext):
    """resnet_layer.py:40-70 (layer shapes from :42-65)."""

    def __init__(self, bottleneck_channel, ext_config=None, bottleneck_transformer=None):
        if not 1 <= int(bottleneck_channel) <= 16:
            raise ValueError("bottleneck_channel %r not supported (1..16)" % (bottleneck_channel,))
        encoder, decoder = _make_encoder_decoder(int(bottleneck_channel))
        from .classifier import Ext4ResNet
        ext = Ext4ResNet(64) if ext_config is not None else None  # resnet_layer.py:66
        super().__init__(encoder=ExtEncoder(encoder, ext, ext_config), decoder=decoder,
                         bottleneck_transformer=bottleneck_transformer)
        self.bottleneck_channel = int(bottleneck_channel)
        # custom/resnet.py:55-60 applies this init to every conv / norm, injected layers included
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def get_mimic_layers(backbone_name, backbone_config, bottleneck_transformer=None):
    """resnet_layer.py:73-87 (same accepted names, same ValueError)."""
    layer1, layer2, layer3, layer4 = None, None, None, None
    backbone_params_config = backbone_config['params']
    layer1_config = backbone_params_config.get('layer1', None)
    if layer1_config is not None:
        layer1_name = layer1_config['name']
        ext_config = backbone_config.get('ext_config', None)
        if layer1_name == 'Bottleneck4SmallResNet' and backbone_name in {'custom_resnet18', 'custom_resnet34'}:
            layer1 = Bottleneck4LargeResNet(layer1_config['bottleneck_channel'], ext_config, bottleneck_transformer)
        elif layer1_name == 'Bottleneck4LargeResNet' \
                and backbone_name in {'custom_resnet50', 'custom_resnet101', 'custom_resnet152'}:
            layer1 = Bottleneck4LargeResNet(layer1_config['bottleneck_channel'], ext_config, bottleneck_transformer)
        else:
            raise ValueError('layer1_name `{}` is not expected'.format(layer1_name))
    return layer1, layer2, layer3, layer4
