"""Model factory / checkpoints -- drop-in mirror of src/models/__init__.py:11-70."""
import os

import torch
from torch import nn

from . import rcnn
from .transformer import get_bottleneck_transformer


def is_main_process():
    import torch.distributed as dist
    return (not dist.is_available()) or (not dist.is_initialized()) or dist.get_rank() == 0


def save_ckpt(model, optimizer, lr_scheduler, best_value, config, args, output_file_path):
    """__init__.py:11-17 (written by rank 0 only, like misc_util.save_on_master)."""
    parent = os.path.dirname(output_file_path)
    if parent:
        os.makedirs(parent, exist_ok=True)
    model_state_dict = \
        model.module.state_dict() if isinstance(model, nn.parallel.DistributedDataParallel) else model.state_dict()
    if is_main_process():
        torch.save({'model': model_state_dict, 'optimizer': optimizer.state_dict(), 'best_value': best_value,
                    'lr_scheduler': lr_scheduler.state_dict(), 'config': config, 'args': args},
                   output_file_path)


def load_ckpt(ckpt_file_path, model=None, optimizer=None, lr_scheduler=None, strict=True):
    """__init__.py:20-35."""
    if ckpt_file_path is None or not os.path.exists(ckpt_file_path):
        print('ckpt file is not found at `{}`'.format(ckpt_file_path))
        return None, None
    ckpt = torch.load(ckpt_file_path, map_location='cpu', weights_only=False)
    if model is not None:
        print('Loading model parameters')
        model.load_state_dict(ckpt['model'], strict=strict)
    if optimizer is not None:
        print('Loading optimizer parameters')
        optimizer.load_state_dict(ckpt['optimizer'])
    if lr_scheduler is not None:
        print('Loading scheduler parameters')
        lr_scheduler.load_state_dict(ckpt['lr_scheduler'])
    return ckpt.get('best_value', 0.0), ckpt['config'], ckpt['args']


def get_model(model_config, device, strict=True, bottleneck_transformer=None):
    """__init__.py:38-57."""
    model_name = model_config['name']
    ckpt_file_path = model_config['ckpt']
    model_params_config = model_config['params']
    if model_name in rcnn.MODEL_CLASS_DICT:
        backbone_config = model_config['backbone']
        if bottleneck_transformer is None and 'bottleneck_transformer' in model_config:
            bottleneck_transformer = get_bottleneck_transformer(model_config['bottleneck_transformer'])
        model = rcnn.get_model(model_name, backbone_config=backbone_config, strict=strict,
                               bottleneck_transformer=bottleneck_transformer, **model_params_config)
        if 'ext_config' in backbone_config:  # __init__.py:49-52: the filter has its own checkpoint
            ext_config = backbone_config['ext_config']
            load_ckpt(ext_config['ckpt'], model=model.get_ext_classifier())
            strict = False
    else:
        raise ValueError('model_name `{}` is not expected'.format(model_name))
    load_ckpt(ckpt_file_path, model=model, strict=strict)
    return model.to(device)


def get_iou_types(model):
    model_without_ddp = model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model
    iou_type_list = ['bbox']
    if isinstance(model_without_ddp, rcnn.MaskRCNN):
        iou_type_list.append('segm')
    if isinstance(model_without_ddp, rcnn.KeypointRCNN):
        iou_type_list.append('keypoints')
    return iou_type_list
