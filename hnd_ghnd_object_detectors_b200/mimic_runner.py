"""`mimic_runner.py` -- drop-in mirror of src/mimic_runner.py (flags :17-29, loop :38-59,
distill :62-106, main :124-151) on the B200 CUDA path.

  python -m hnd_ghnd_object_detectors_b200.mimic_runner --config X.yaml -distill [--json '{...}']
  torchrun --nproc-per-node 8 -m hnd_ghnd_object_detectors_b200.mimic_runner --config X.yaml -distill

What is the same: YAML schema (with !join), --json deep-merge, teacher/student construction,
frozen_modules, Adam + MultiStepLR + epoch-0 warm-up, best-checkpoint format and keys.
What differs: DistributedDataParallel is replaced by one flat-buffer NCCL all-reduce + fused Adam
(parallel.py; rank 0's parameters and BN buffers are broadcast at start like DDP's constructor does);
COCO loading and COCO mAP (pycocotools; SURVEY.md "out of scope") are not re-implemented --
`dataset.name: synthetic` gives random images so both entry points run anywhere: `-distill` trains,
and the evaluation entry point (mimic_runner.py:109-121,148-151) runs the full detector forward of the
teacher and of the student (with the quantize/dequantize splice under -transform_bottleneck) and
reports model time, detections and the bottleneck's bytes on the wire instead of mAP.
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
        if optimizer.flat is None:
            optimizer.attach(distillation_box.flat)
        parallel.allreduce_flat_grad(distillation_box.flat)  # p.grad are views of this buffer
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
    flat = distillation_box.flatten_parameters()
    if parallel.world_size() > 1:
        # DistributedDataParallel's constructor broadcasts rank 0's parameters and buffers
        parallel.broadcast_flat_params(flat)
        parallel.broadcast_buffers(student_model)
    ckpt_file_path = config['student_model']['ckpt']
    optim_config = train_config['optimizer']
    if optim_config['type'].lower() != 'adam':
        raise ValueError("optimizer `{}`: the fused CUDA path implements Adam".format(optim_config['type']))
    optimizer = FusedAdam([p for p in student_model.parameters() if p.requires_grad],
                          grad_scale=1.0 / parallel.world_size(), flat=flat, **optim_config['params'])
    scheduler_config = train_config['scheduler']
    scheduler_cls = getattr(torch.optim.lr_scheduler, scheduler_config['type'])
    lr_scheduler = scheduler_cls(optimizer, **scheduler_config['params'])
    best_val = 0.0
    import os
    if os.path.exists(ckpt_file_path):
        best_val, _, _ = load_ckpt(ckpt_file_path, optimizer=optimizer, lr_scheduler=lr_scheduler)
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
        # COCO mAP (main_util.evaluate + pycocotools) is outside the hot path: the latest student is
        # kept instead of the best-mAP one
        save_ckpt(student_model, optimizer, lr_scheduler, best_val, config, args, ckpt_file_path)
        lr_scheduler.step()
    if distributed:
        dist.barrier()
    total_time = time.time() - start_time
    print('Training time {}'.format(str(datetime.timedelta(seconds=int(total_time)))))
    return distillation_box


def evaluate_model(model, data_loader, device, max_batches=None):
    """main_util.evaluate (src/utils/main_util.py:76-113) without the COCO scorer: eval-mode forward of
    the full detector per batch, model time by CUDA-synchronised wall clock like the reference,
    detections moved to the host.  Returns a dict of averaged stats."""
    from . import file_util
    model.eval()
    n_img, n_det, t_model, wire = 0, 0, 0.0, []
    layer1 = model.backbone.body.layer1
    log_wire = getattr(layer1, 'use_bottleneck_transformer', False) and \
        getattr(layer1, 'bottleneck_transformer', None) is not None
    if log_wire:
        layer1.data_logging = True
    with torch.no_grad():
        for it, (images, targets) in enumerate(data_loader):
            if max_batches is not None and it >= max_batches:
                break
            images = [img.to(device) for img in images]
            torch.cuda.synchronize()
            t0 = time.time()
            outputs = model(images)
            outputs = [{k: v.to('cpu') for k, v in t.items()} for t in outputs]
            t_model += time.time() - t0
            n_img += len(images)
            n_det += sum(len(o['boxes']) for o in outputs)
            if log_wire and getattr(layer1, 'last_bottleneck', None) is not None:
                wire.append(file_util.get_binary_object_size(layer1.last_bottleneck) / len(images))
    if log_wire:
        layer1.data_logging = False
    stats = {'images': n_img, 'detections': n_det, 'model_time': t_model / max(n_img, 1)}
    if wire:
        stats['bottleneck_kb_per_image'] = sum(wire) / len(wire)
    print('Averaged stats: model_time: {:.4f} s/img  detections/img: {:.1f}{}'.format(
        stats['model_time'], n_det / max(n_img, 1),
        '  bottleneck: {:.1f} KB/img'.format(stats['bottleneck_kb_per_image']) if wire else ''))
    return stats


def evaluate(teacher_model, student_model, test_data_loader, device, student_only, use_bottleneck_transformer):
    """mimic_runner.py:109-121."""
    teacher_model.distill_backbone_only = False
    student_model.distill_backbone_only = False
    student_model.backbone.body.layer1.use_bottleneck_transformer = use_bottleneck_transformer
    results = {}
    if not student_only:
        print('[Teacher model]')
        results['teacher'] = evaluate_model(teacher_model, test_data_loader, device)
    print('\n[Student model]')
    results['student'] = evaluate_model(student_model, test_data_loader, device)
    return results


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
    results = evaluate(teacher_model, student_model, val_loader, device, args.skip_teacher_eval,
                       args.transform_bottleneck)
    if distributed:
        dist.destroy_process_group()
    return results


if __name__ == '__main__':
    parser = get_argparser()
    main(parser.parse_args())
