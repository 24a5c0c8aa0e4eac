"""Wire-format size of the bottleneck -- mirror of src/myutils/common/file_util.py:52-53.

Split computing ships the student's bottleneck over the network as a pickled object: a fp32 tensor, its
fp16 copy, or the `QuantizedTensor(tensor: uint8, scale, zero_point)` produced by the 8-bit quantizer.
`get_binary_object_size` is the reference's measure of those bytes (sys.getsizeof of the pickle, in
KB); `bottleneck_wire_sizes` is what its DataLogger (src/structure/transformer.py:76-91) records per
bottleneck: the size as-is, as fp16 and 8-bit quantized."""
import pickle
import sys

import torch


def _to_host(x):
    """The wire carries host bytes: move tensors (also inside QuantizedTensor-like tuples) to the CPU."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu()
    if isinstance(x, tuple) and hasattr(x, '_fields'):
        return type(x)(*[_to_host(v) for v in x])
    if isinstance(x, (tuple, list)):
        return type(x)(_to_host(v) for v in x)
    return x


def get_binary_object_size(x, unit_size=1024):
    return sys.getsizeof(pickle.dumps(_to_host(x))) / unit_size


def bottleneck_wire_sizes(z, num_bits=8):
    """(data_size, fp16_data_size, quantized_data_size) in KB, as DataLogger.__call__ appends them."""
    from . import tensor_util
    if z is None:
        return 0.0, 0.0, 0.0
    data_size = get_binary_object_size(z)
    if not isinstance(z, torch.Tensor):
        return data_size, None, None
    fp16 = get_binary_object_size(z.short())
    quantized = get_binary_object_size(tensor_util.quantize_tensor(z, num_bits=num_bits))
    return data_size, fp16, quantized
