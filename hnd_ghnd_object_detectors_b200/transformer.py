"""Bottleneck transformers -- drop-in mirror of src/structure/transformer.py:22-29,131-174
(Compose, Quantizer, Dequantizer, registry, get_bottleneck_transformer).  The JPEG codec and the
DataLogger of the reference are analysis tools outside the hot path."""
from . import tensor_util


class Compose(object):
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, image, target):
        for t in self.transforms:
            image, target = t(image, target)
        return image, target


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
