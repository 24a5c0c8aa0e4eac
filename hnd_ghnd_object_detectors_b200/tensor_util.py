"""8-bit bottleneck quantizer -- drop-in mirror of src/myutils/pytorch/tensor_util.py.

Same names and return types: quantize_tensor(x, num_bits=8) -> QuantizedTensor(tensor: uint8 on
x's device, scale: 0-dim fp32 tensor, zero_point: python int); dequantize_tensor(q_x) -> fp32.
The arithmetic is the fused CUDA kernel pair behind ghnd_quantize_u8 / ghnd_dequantize_u8
(bit-exact to the reference, see include/ghnd_b200.h); CPU tensors are rejected (no fallback).
"""
import os
from collections import namedtuple

import torch

from . import _lib, ops

QuantizedTensor = namedtuple('QuantizedTensor', ['tensor', 'scale', 'zero_point'])

_NAN_MARKER = -2 ** 31
_BARRIER_TIMEOUT_MARKER = -2 ** 31 + 1  # quant_fused_kernel: a CTA never reached the grid barrier


def check_zero_point(zero_point):
    if zero_point == _NAN_MARKER:
        raise ValueError("cannot convert float NaN to integer")  # int(nan) in tensor_util.py:15
    if zero_point == _BARRIER_TIMEOUT_MARKER:
        raise _lib.GhndError("ghnd_quantize_u8: grid barrier timed out (workspace not zeroed, or the "
                             "device cannot co-schedule one CTA per SM)")
    return zero_point


def _scale_mode():
    # "div": scale=(max-min)/255 by IEEE division -- what the reference computes on CPU and what the
    # oracle / goldens pin.  "recip": (max-min)*fl(1/255), what torch computes for the same Python
    # line on a CUDA device.  See DESIGN.md "quantizer".
    return _lib.QSCALE_RECIP if os.environ.get("GHND_QUANT_SCALE", "div") == "recip" else _lib.QSCALE_DIV


def quantize_tensor(x, num_bits=8):
    """tensor_util.py:8-18."""
    if x.numel() == 0:
        raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0")
    q, qp = ops.quantize_u8(x, num_bits, _scale_mode())
    host = qp.cpu()  # the reference API returns a python int zero-point: one unavoidable D2H sync
    zero_point = int(host[1])
    check_zero_point(zero_point)
    scale = qp[0:1].view(torch.float32).reshape(())
    return QuantizedTensor(tensor=q, scale=scale, zero_point=zero_point)


def dequantize_tensor(q_x):
    """tensor_util.py:21-22."""
    t = q_x.tensor
    if not t.is_cuda:
        raise _lib.GhndError("dequantize_tensor needs a CUDA tensor (no CPU fallback)")
    scale = q_x.scale
    scale = float(scale.item()) if isinstance(scale, torch.Tensor) else float(scale)
    qp = ops.make_qparams(scale, q_x.zero_point, t.device)
    return ops.dequantize_u8(t, qp)
