"""`mimic_runner.py` -- drop-in mirror of src/mimic_runner.py (flags :17-29, loop :38-59,
distill :62-106, main :124-151) on the B200 CUDA path.

  python -m hnd_ghnd_object_detectors_b200.mimic_runner --config X.yaml -distill [--json '{...}']
  torchrun --nproc-per-node 8 -m hnd_ghnd_object_detectors_b200.mimic_runner --config X.yaml -distill

What is the same: YAML schema (with !join), --json deep-merge, teacher/student construction,
frozen_modules, Adam + MultiStepLR + epoch-0 warm-up, best-checkpoint format and keys.
What differs: DistributedDataParallel is replaced by one flat-buffer NCCL all-reduce + fused Adam
(parallel.py); COCO loading and COCO evaluation (pycocotools; SURVEY.md "out of scope") are used
only if available -- `dataset.name: synthetic` gives random images so the loop runs anywhere.
"""
import argparse
import datetime
import random
import time

import torch
from torch import distributed as dist

from .prefetch import DevicePrefetcher
from . import main_util, module_util, parallel, yaml_util
from .models import get_model, load_ckpt, save_ckpt
from .optim import FusedAdam
from .tool import DistillationBox


def get_argparser():
    argparser = argparse.ArgumentParser(description='Mimic Runner')
    argparser.add_argument('--config', required=True, help='yaml file path')
    argparser.add_argument('--device', default='cuda', help='device')
    argparser.add_argument('--json', help='dictionary to overwrite config')
    argparser.add_argument('-distill', action='store_true', help='distill a teacher model')
    argparser.add_argument('-skip_teacher_eval', action='store_true', help='skip teacher model evaluation in testing')
    argparser.add_argument('-transform_bottleneck', action='store_true',
                           help='use bottleneck transformer (if defined in yaml) in testing')
    # distributed training parameters
    argparser.add_argument('--world_size', default=1, type=int, help='number of distributed processes')
    argparser.add_argument('--dist_url', default='env://', help='url used to set up distributed training')
    return argparser


def freeze_modules(student_model, student_model_config):
    for student_path in student_model_config['frozen_modules']:
        student_module = module_util.get_module(student_model, student_path)
        module_util.freeze_module_params(student_module)


class SyntheticLoader(object):
    """Random images of one fixed size, sharded over ranks like DistributedSampler."""

    def __init__(self, dataset_config, batch_size, device):
        self.n = int(dataset_config.get('num_samples', 64))
        self.h, self.w = dataset_config.get('image_size', [800, 1333])
        self.batch_size, self.device = batch_size, device
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.indices = parallel.shard_indices(self.n, rank, parallel.world_size())

    def __len__(self):
        return max(len(self.indices) // self.batch_size, 1)

    def __iter__(self):
        for b in range(len(self)):
            idx = self.indices[b * self.batch_size:(b + 1) * self.batch_size]
            images, targets = [], []
            for i in idx:
                g = torch.Generator().manual_seed(i)
                images.append(torch.rand(3, self.h, self.w, generator=g))
                targets.append({'boxes': torch.tensor([[10., 10., 100., 100.]]), 'labels': torch.tensor([1])})
            yield images, targets


def get_data_loaders(dataset_config, batch_size, device):
    if dataset_config.get('name') == 'synthetic':
        loader = SyntheticLoader(dataset_config, batch_size, device)
        return loader, loader
    raise RuntimeError("COCO loading needs pycocotools and the dataset on disk (outside the B200 hot "
                       "path); use dataset.name: synthetic or plug your own DataLoader into distill()")


def distill_model(distillation_box, data_loader, optimizer, log_freq, device, epoch):
    """mimic_runner.py:38-59."""
    lr_scheduler = None
    if epoch == 0:
        warmup_factor = 1.0 / 1000.0
        warmup_iters = min(1000, len(data_loader) - 1)
        if warmup_iters > 0:
            lr_scheduler = main_util.warmup_lr_scheduler(optimizer, warmup_iters, warmup_factor)
    t0, seen, last = time.time(), 0, None
    # the copy of batch i+1 overlaps step i on a side stream (prefetch.py)
    for it, (images, targets) in enumerate(DevicePrefetcher(data_loader, device)):
        loss = distillation_box(images, targets)
        optimizer.zero_grad()
        loss.backward()
        if distillation_box.flat is not None and optimizer.flat is None:
            optimizer.attach(distillation_box.flat)
        parallel.allreduce_flat_grad(distillation_box.flat)
        optimizer.step()
        if lr_scheduler is not None:
            lr_scheduler.step()
        seen += len(images)
        if it % log_freq == 0:
            last = loss.item()  # the only host sync of the loop (reference: every step)
            print('Epoch: [{}] [{}/{}] loss: {:.6g} lr: {:.6f} img/s/rank: {:.1f}'.format(
                epoch, it, len(data_loader), last, optimizer.param_groups[0]['lr'],
                seen / max(time.time() - t0, 1e-9)))
    return last


def distill(teacher_model, student_model, train_data_loader, val_data_loader, device, distributed,
            distill_backbone_only, config, args):
    """mimic_runner.py:62-106."""
    train_config = config['train']
    distillation_box = DistillationBox(teacher_model, student_model, train_config['criterion'])
    ckpt_file_path = config['student_model']['ckpt']
    optim_config = train_config['optimizer']
    if optim_config['type'].lower() != 'adam':
        raise ValueError("optimizer `{}`: the fused CUDA path implements Adam".format(optim_config['type']))
    optimizer = FusedAdam([p for p in student_model.parameters() if p.requires_grad],
                          grad_scale=1.0 / parallel.world_size(), **optim_config['params'])
    scheduler_config = train_config['scheduler']
    scheduler_cls = getattr(torch.optim.lr_scheduler, scheduler_config['type'])
    lr_scheduler = scheduler_cls(optimizer, **scheduler_config['params'])
    best_val = 0.0
    import os
    if os.path.exists(ckpt_file_path):
        best_val, _, _ = load_ckpt(ckpt_file_path, optimizer=None, lr_scheduler=lr_scheduler)
    start_time = time.time()
    for epoch in range(train_config['num_epochs']):
        teacher_model.eval()
        student_model.train()
        teacher_model.distill_backbone_only = distill_backbone_only
        student_model.distill_backbone_only = distill_backbone_only
        student_model.backbone.body.layer1.use_bottleneck_transformer = False
        last = distill_model(distillation_box, train_data_loader, optimizer, train_config['log_freq'],
                             device, epoch)
        student_model.distill_backbone_only = False
        student_model.backbone.body.layer1.use_bottleneck_transformer = args.transform_bottleneck
        # COCO mAP evaluation (main_util.evaluate) is outside the hot path: keep the latest student
        save_ckpt(student_model, optimizer, lr_scheduler, best_val, config, args, ckpt_file_path)
        lr_scheduler.step()
    if distributed:
        dist.barrier()
    total_time = time.time() - start_time
    print('Training time {}'.format(str(datetime.timedelta(seconds=int(total_time)))))
    return distillation_box


def main(args):
    config = yaml_util.load_yaml_file(args.config)
    if args.json is not None:
        main_util.overwrite_config(config, args.json)
    distributed, device_ids = main_util.init_distributed_mode(args.world_size, args.dist_url)
    if not torch.cuda.is_available():
        raise SystemExit("mimic_runner needs a CUDA device: the B200 path has no CPU fallback")
    device = torch.device('cuda', device_ids[0] if device_ids else torch.cuda.current_device())
    random.seed(0)
    teacher_model = get_model(config['teacher_model'], device)
    module_util.freeze_module_params(teacher_model)
    student_model_config = config['student_model']
    student_model = get_model(student_model_config, device)
    freeze_modules(student_model, student_model_config)
    print('Updatable parameters: {}'.format(module_util.get_updatable_param_names(student_model)))
    distill_backbone_only = student_model_config['distill_backbone_only']
    train_config = config['train']
    train_loader, val_loader = get_data_loaders(config['dataset'], train_config['batch_size'], device)
    if args.distill:
        distill(teacher_model, student_model, train_loader, val_loader, device, distributed,
                distill_backbone_only, config, args)
        load_ckpt(config['student_model']['ckpt'], model=student_model)
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    parser = get_argparser()
    main(parser.parse_args())
