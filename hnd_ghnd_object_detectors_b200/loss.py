"""HND / GHND criterion -- drop-in mirror of src/distillation/loss.py (CustomLoss,
GeneralizedCustomLoss, LOSS_DICT, get_loss) with the sub-criterion lookup of
src/myutils/pytorch/func_util.py:9-13 restricted to what the configs use: MSELoss(reduction='sum').
The arithmetic is the fused SSE forward+backward kernel (ghnd_sse_fwd_bwd)."""
import torch
from torch import nn

from . import _lib, ops


class _SseFunction(torch.autograd.Function):
    """factor * sum((teacher - student)^2) with d/dstudent = 2*factor*(student - teacher)."""

    @staticmethod
    def forward(ctx, teacher, student, factor):
        if not student.is_cuda:
            raise _lib.GhndError("GHND loss runs on CUDA only (no CPU fallback)")
        t16 = teacher.detach().to(torch.float16).contiguous()
        s16 = student.detach().to(torch.float16).contiguous()
        n = s16.numel()
        pad = (-n) % 8
        if pad:  # the kernel works on 16-byte vectors
            t16 = torch.cat([t16.reshape(-1), t16.new_zeros(pad)])
            s16 = torch.cat([s16.reshape(-1), s16.new_zeros(pad)])
        grad = torch.empty(s16.shape, dtype=torch.bfloat16, device=s16.device)
        out = ops.sse_fwd_bwd([(t16, s16, grad, factor, False)])
        ctx.save_for_backward(grad)
        ctx.n, ctx.shape, ctx.dtype = n, student.shape, student.dtype
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        gs = grad.reshape(-1)[:ctx.n].reshape(ctx.shape).to(ctx.dtype) * g
        return None, gs, None


class SumSquaredError(nn.Module):
    """nn.MSELoss(reduction='sum') on the CUDA kernel; called as criterion(teacher_out, student_out)
    like loss.py:29 (the student is the `target` argument there; the gradient still reaches it)."""

    def forward(self, teacher_output, student_output):
        return _SseFunction.apply(teacher_output, student_output, 1.0)


def get_sub_criterion(loss_type, param_dict):
    if loss_type.lower() != 'mseloss' or param_dict.get('reduction', 'mean') != 'sum':
        raise ValueError("criterion `{}` with params {} is not available on the CUDA path "
                         "(the HND/GHND configs use MSELoss(reduction='sum'))".format(loss_type, param_dict))
    return SumSquaredError()


class CustomLoss(nn.Module):
    def __init__(self, criterion_config):
        super().__init__()
        self.org_loss_factor = criterion_config['params']['org_loss_factor']
        term_dict = dict()
        for loss_name, loss_config in criterion_config['terms'].items():
            sub_criterion_config = loss_config['criterion']
            sub_criterion = get_sub_criterion(sub_criterion_config['type'], sub_criterion_config['params'])
            term_dict[loss_name] = (loss_config['ts_modules'], sub_criterion, loss_config['factor'])
        self.term_dict = term_dict

    def forward(self, *args, **kwargs):
        raise NotImplementedError('forward function is not implemented')


class GeneralizedCustomLoss(CustomLoss):
    def __init__(self, criterion_config):
        super().__init__(criterion_config)

    def forward(self, output_dict, org_loss_dict):
        loss_dict = dict()
        for loss_name, ((teacher_path, teacher_output), (student_path, student_output)) in output_dict.items():
            _, criterion, factor = self.term_dict[loss_name]
            loss_dict[loss_name] = criterion(teacher_output, student_output) * factor
        sub_total_loss = sum(loss for loss in loss_dict.values())
        if self.org_loss_factor == 0:
            return sub_total_loss
        return sub_total_loss + self.org_loss_factor * sum(loss for loss in org_loss_dict.values())


LOSS_DICT = {
    'general': GeneralizedCustomLoss
}


def get_loss(criterion_config):
    criterion_type = criterion_config['type']
    if criterion_type in LOSS_DICT:
        return LOSS_DICT[criterion_type](criterion_config)
    raise ValueError('criterion_type `{}` is not expected'.format(criterion_type))
