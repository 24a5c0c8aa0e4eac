"""Neural-filter head -- drop-in mirror of src/models/ext/classifier.py (BaseExtClassifier :8-13,
Ext4ResNet :16-37, get_ext_classifier :40-43).

Same class name, constructor argument, submodule layout and state_dict keys (`extractor.{1,2,4,5,7,8}.*`,
`linear.*`) so the filter checkpoints written by the reference's ext_runner.py load with strict=True.
The nn modules only HOLD the parameters; inference runs the hand-written kernels of csrc/ext_filter.cu:
one HBM-bound adaptive average pool over the 64-channel stem output, then three small fp32 convolutions
with the eval-mode BatchNorm folded in, the 8x8 pool + Linear + softmax.  Training the filter
(ext_runner.py, cross-entropy on image-level labels) is outside the B200 hot path and raises."""
import torch
from torch import nn

from . import _lib, ops


class BaseExtClassifier(nn.Module):
    def __init__(self, ext_idx):
        super().__init__()
        self.ext_idx = ext_idx

    def forward(self, *args):
        raise NotImplementedError('forward function is not implemented')


def _fold(conv, bn):
    """conv bias + eval BatchNorm -> per-channel (scale, shift) and weights repacked [R][S][C][K]."""
    sc = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros_like(sc)
    sh = bn.bias.detach().float() + (bias - bn.running_mean.detach().float()) * sc
    w = conv.weight.detach().float().permute(2, 3, 1, 0).contiguous()
    return w, sc.contiguous(), sh.contiguous()


class Ext4ResNet(BaseExtClassifier):
    def __init__(self, input_channel):
        super().__init__(ext_idx=0)
        self.extractor = nn.Sequential(
            nn.AdaptiveAvgPool2d((64, 64)),
            nn.Conv2d(input_channel, 64, kernel_size=4, stride=2),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
            nn.Conv2d(64, 32, kernel_size=3, stride=2),
            nn.BatchNorm2d(32),
            nn.ReLU(inplace=True),
            nn.Conv2d(32, 16, kernel_size=2, stride=1),
            nn.BatchNorm2d(16),
            nn.ReLU(inplace=True),
            nn.AdaptiveAvgPool2d((8, 8))
        )
        self.linear = nn.Linear(16 * 8 * 8, 2)

    def forward_nhwc16(self, x):
        """x: NHWC 16-bit stem output [N,H,W,C] (what the encode / body plans hold) -> [N,2] softmax."""
        if self.training:
            raise _lib.GhndError("training the neural filter (ext_runner.py) is outside the B200 hot path; "
                                 "call .eval()")
        if not x.is_cuda:
            raise _lib.GhndError("Ext4ResNet runs on CUDA only (no CPU fallback)")
        e = self.extractor
        y = ops.adaptive_avgpool_nhwc16(x, 64, 64)
        for conv, bn in ((e[1], e[2]), (e[4], e[5]), (e[7], e[8])):
            w, sc, sh = _fold(conv, bn)
            y = ops.small_conv_f32(y, w, sc, sh, True, conv.stride[0])
        return ops.avgpool_linear(y, 8, 8, self.linear.weight.detach().float().contiguous(),
                                  self.linear.bias.detach().float().contiguous(), softmax=True)

    def forward(self, x):
        """x: NCHW fp32 [N,C,H,W] like the reference module; eval: softmax probabilities (classifier.py:37)."""
        if not x.is_cuda:
            raise _lib.GhndError("Ext4ResNet runs on CUDA only (no CPU fallback)")
        return self.forward_nhwc16(ops.to_nhwc16(x))


def get_ext_classifier(backbone):
    from torchvision.models.resnet import ResNet
    if isinstance(backbone, ResNet):
        return Ext4ResNet(64)
    raise ValueError('type of backbone `{}` is not expected'.format(type(backbone)))
