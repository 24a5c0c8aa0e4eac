"""Bottleneck transformers -- drop-in mirror of src/structure/transformer.py:22-29,131-174
(Compose, Quantizer, Dequantizer, DataLogger :58-91, registry, get_bottleneck_transformer).  The JPEG
codec of the reference is an analysis tool outside the hot path."""
from . import file_util, tensor_util


class Compose(object):
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, image, target):
        for t in self.transforms:
            image, target = t(image, target)
        return image, target


class DataLogger(object):
    """transformer.py:58-91: records the bytes the bottleneck would occupy on the wire (pickled as-is,
    as fp16 and 8-bit quantized) plus its C/H/W shape; passes z through unchanged."""

    def __init__(self, num_bits=8):
        self.num_bits4quant = num_bits
        self.data_size_list = list()
        self.fp16_data_size_list = list()
        self.quantized_data_size_list = list()
        self.tensor_shape_list = list()

    def get_data(self):
        return self.data_size_list.copy(), self.fp16_data_size_list, \
            self.quantized_data_size_list.copy(), self.tensor_shape_list.copy()

    def clear(self):
        self.data_size_list.clear()
        self.fp16_data_size_list.clear()
        self.quantized_data_size_list.clear()
        self.tensor_shape_list.clear()

    def __call__(self, z, target):
        data_size, fp16_data_size, quantized_data_size = file_util.bottleneck_wire_sizes(z, self.num_bits4quant)
        self.data_size_list.append(data_size)
        self.fp16_data_size_list.append(fp16_data_size)
        self.quantized_data_size_list.append(quantized_data_size)
        self.tensor_shape_list.append([0, 0, 0] if z is None else [z.shape[1], z.shape[2], z.shape[3]])
        return z, target


class Quantizer(object):
    def __init__(self, num_bits=8):
        self.num_bits = num_bits

    def __call__(self, z, target):
        if self.num_bits == 16:
            return z.half(), target
        qz = tensor_util.quantize_tensor(z, num_bits=self.num_bits)
        return qz, target


class Dequantizer(object):
    def __init__(self, num_bits=8):
        # num_bits should be the same as Quantizer
        self.num_bits = num_bits

    def __call__(self, qz, target):
        if self.num_bits == 16:
            return qz.float(), target
        z = tensor_util.dequantize_tensor(qz)
        return z, target


TRANSFORMER_CLASS_DICT = {
    'quantizer': Quantizer,
    'dequantizer': Dequantizer
}


def get_bottleneck_transformer(transformer_config):
    component_list = list()
    components_config = transformer_config['components']
    for component_name in transformer_config['order']:
        param_config = components_config[component_name]['params']
        if component_name not in TRANSFORMER_CLASS_DICT:
            raise KeyError('transformer `{}` is not expected'.format(component_name))
        obj_class = TRANSFORMER_CLASS_DICT[component_name]
        component_list.append(obj_class(**param_config))
    return Compose(component_list) if len(component_list) > 0 else None
