"""Fused Adam over the student's flat parameter buffer (torch.optim.Adam semantics,
src/mimic_runner.py:52-54 / func_util.get_optimizer).  One kernel per step instead of 25
per-tensor update loops; when the parameters were flattened by DistillationBox the gradients are
consumed in place (no gather)."""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 flat=None, grad_scale=1.0):
        if amsgrad:
            raise ValueError("amsgrad is not supported by the fused kernel")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat = flat
        self.grad_scale = grad_scale
        self._state_flat = None
        self._step = 0

    def attach(self, flat):
        """Use a FlatParams (engine.py) buffer: update all tensors with a single kernel."""
        self.flat = flat

    @torch.no_grad()
    def step(self, closure=None):
        self._step += 1
        group = self.param_groups[0]
        lr, (b1, b2), eps, wd = group['lr'], group['betas'], group['eps'], group['weight_decay']
        if self.flat is not None:
            f = self.flat
            if self._state_flat is None:
                self._state_flat = (torch.zeros_like(f.flat), torch.zeros_like(f.flat))
            for n in f.names:  # gradients written elsewhere (e.g. by DDP) are gathered back
                p = f.params[n]
                if p.grad is not None and p.grad.data_ptr() != f.grads[n].data_ptr():
                    f.grads[n].copy_(p.grad)
            m, v = self._state_flat
            ops.adam_step(f.flat, f.grad, m, v, lr, b1, b2, eps, wd, self.grad_scale, self._step)
            return None
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                g = p.grad.contiguous()
                ops.adam_step(p.data, g, st['exp_avg'], st['exp_avg_sq'], group['lr'], group['betas'][0],
                              group['betas'][1], group['eps'], group['weight_decay'], self.grad_scale,
                              self._step)
        return None
