"""DistillationBox -- drop-in mirror of src/distillation/tool.py.

Same constructor and call signature: DistillationBox(teacher, student, criterion_config)(images,
targets) -> scalar loss with a grad_fn; `loss.backward()` deposits the gradients of the trainable
student tensors.  The reference's forward-hook plumbing (tool.py:19-35) is replaced by the fused
GhndPlan (engine.py): `ts_modules` paths must name `backbone.body.layerK` on both models (what every
HND/GHND config uses) and keep their pairing semantics; the whole step -- teacher forward, student
forward, multi-level SSE, student backward -- runs as one fixed kernel sequence (CUDA graph).
"""
import random

import torch
from torch import nn
from torch.nn import DataParallel
from torch.nn.parallel.distributed import DistributedDataParallel

from . import _lib
from .engine import LEVELS, FlatParams, GhndPlan
from .loss import get_loss
from .rcnn import KeypointRCNN, round_up


class _GradInjector(torch.autograd.Function):
    """Hands the gradients the fused backward already produced to autograd."""

    @staticmethod
    def forward(ctx, loss, box, *params):
        ctx.box = box
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        # The fused backward wrote every gradient into the flat buffer during forward(); here they
        # are scaled by the incoming gradient with ONE in-place multiply and each `p.grad` is made a
        # VIEW of that buffer, so the data-parallel all-reduce of `flat.grad` and the fused Adam see
        # exactly what autograd users see (no per-tensor copies, no 25 multiplies).  Gradients are
        # assigned, not accumulated: the plan overwrites the buffer every step.
        flat = ctx.box.flat
        flat.grad.mul_(g)
        for p_name in flat.names:
            flat.params[p_name].grad = flat.grads[p_name]
        return (None, None) + (None,) * len(flat.names)


def _level_of(path):
    prefix = 'backbone.body.'
    if not path.startswith(prefix) or path[len(prefix):] not in LEVELS:
        raise ValueError("ts_modules path `{}` is not a backbone.body.layerK module; the fused CUDA "
                         "distillation path supports the HND/GHND configs' layer1..layer4 taps".format(path))
    return path[len(prefix):]


class DistillationBox(nn.Module):
    def __init__(self, teacher_model, student_model, criterion_config, use_cuda_graph=True):
        super().__init__()
        self.teacher_model = teacher_model
        self.student_model = student_model
        self.target_module_pairs = list()
        self.levels, self.factors = [], {}
        for loss_name, loss_config in criterion_config['terms'].items():
            teacher_path, student_path = loss_config['ts_modules']
            self.target_module_pairs.append((teacher_path, student_path))
            lt, ls = _level_of(teacher_path), _level_of(student_path)
            if lt != ls:
                raise ValueError("teacher/student taps must be the same layer ({} vs {})".format(lt, ls))
            self.levels.append(ls)
            self.factors[ls] = float(loss_config['factor'])
        self.criterion = get_loss(criterion_config)  # validates the term types (MSELoss sum)
        if self.criterion.org_loss_factor != 0:
            raise ValueError("org_loss_factor != 0 needs the detection losses (FPN/RPN/RoI heads), "
                             "which are outside the B200 hot path")
        student = self._unwrap(student_model)
        self.require_adjustment = isinstance(student, KeypointRCNN)
        teacher_min = self._unwrap(teacher_model).transform.min_size
        teacher_min = teacher_min if isinstance(teacher_min, (list, tuple)) else (teacher_min,)
        self.use_cuda_graph = use_cuda_graph
        # Resident shapes (one GhndPlan + CUDA graph each, a few GB out of 180 GB): real COCO batches
        # pad to a handful of shapes and the Keypoint path draws from six scales, so keep an LRU of
        # several plans for every model type instead of rebuilding on each shape change.
        self.max_resident_plans = max(len(teacher_min), 4)
        self.flat = None
        self._plans = {}
        self.last_terms = None

    @staticmethod
    def _unwrap(model):
        return model.module if isinstance(model, (DataParallel, DistributedDataParallel)) else model

    def flatten_parameters(self):
        """Move the trainable student tensors into ONE flat fp32 buffer (engine.FlatParams): a single
        all-reduce and a single fused Adam kernel cover them.  Called lazily by the first forward; the
        runner calls it up front to broadcast rank 0's parameters and attach the optimizer."""
        if self.flat is None:
            student = self._unwrap(self.student_model)
            named = [("backbone.body." + k, p) for k, p in student.backbone.body.named_parameters()]
            self.flat = FlatParams(named)
        return self.flat

    def _plan(self, n, hp, wp):
        key = (n, hp, wp)
        plan = self._plans.get(key)
        if plan is None:
            teacher, student = self._unwrap(self.teacher_model), self._unwrap(self.student_model)
            self.flatten_parameters()
            plan = GhndPlan(teacher.backbone.body, student.backbone.body, n, hp, wp, levels=self.levels,
                            factors=self.factors, flat=self.flat,
                            image_mean=student.transform.image_mean, image_std=student.transform.image_std)
            if self.use_cuda_graph:
                plan.capture()
            while len(self._plans) >= self.max_resident_plans:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = plan
        else:
            self._plans[key] = self._plans.pop(key)  # most recently used last
        return plan

    def forward(self, images, targets):
        teacher, student = self._unwrap(self.teacher_model), self._unwrap(self.student_model)
        if student.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        if not images[0].is_cuda:
            raise _lib.GhndError("DistillationBox runs on CUDA only (no CPU fallback)")
        fixed_sizes = None
        if self.require_adjustment:  # tool.py:45-48
            fixed_sizes = [random.choice(teacher.transform.min_size) for _ in images]
        imgs = student._scaled_images(images, fixed_sizes)
        hp = round_up(max(i.shape[1] for i in imgs), 32)
        wp = round_up(max(i.shape[2] for i in imgs), 32)
        plan = self._plan(len(imgs), hp, wp)
        # stale views of the flat gradient buffer (left by the previous backward) are dropped before
        # the plan rewrites it, so a later zero_grad(set_to_none=False) cannot wipe the new gradients
        params = [self.flat.params[n] for n in self.flat.names]
        for n, p in zip(self.flat.names, params):
            if p.grad is not None and p.grad.data_ptr() == self.flat.grads[n].data_ptr():
                p.grad = None
        out = plan.step(imgs)
        self.last_terms = out
        return _GradInjector.apply(out[0], self, *params)
