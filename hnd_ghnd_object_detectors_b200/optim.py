"""Fused Adam over the student's flat parameter buffer (torch.optim.Adam semantics,
src/mimic_runner.py:52-54 / func_util.get_optimizer).  One kernel per step instead of 25
per-tensor update loops; when the parameters were flattened by DistillationBox the gradients are
consumed in place (`p.grad` is a view of the flat gradient buffer, tool._GradInjector), so whatever
an all-reduce left in that buffer is what the update uses.

state_dict()/load_state_dict() speak torch.optim.Adam's format (per-parameter `step`, `exp_avg`,
`exp_avg_sq`), so a checkpoint written here resumes here or under torch.optim.Adam and vice versa
(src/models/__init__.py:11-35 stores/restores `optimizer.state_dict()`)."""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 flat=None, grad_scale=1.0):
        if amsgrad:
            raise ValueError("amsgrad is not supported by the fused kernel")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat = None
        self.grad_scale = grad_scale
        self._state_flat = None
        self._step = 0
        self._pending_state = None  # a loaded per-parameter state waiting for the flat buffer
        if flat is not None:
            self.attach(flat)

    # ---- flat buffer --------------------------------------------------------------------------
    def attach(self, flat):
        """Use a FlatParams (engine.py) buffer: update all tensors with a single kernel.  Moments
        accumulated per tensor so far (or loaded from a checkpoint) move into the flat buffers."""
        if self.flat is flat:
            return
        self.flat = flat
        m, v = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat)
        self._state_flat = (m, v)
        self._m_views, self._v_views = {}, {}
        off = 0
        for n in flat.names:
            p = flat.params[n]
            self._m_views[n] = m[off:off + p.numel()].view(p.shape)
            self._v_views[n] = v[off:off + p.numel()].view(p.shape)
            off += ((p.numel() + 3) // 4) * 4
        for n in flat.names:  # adopt existing per-tensor state (eager steps before the first plan)
            st = self.state.get(flat.params[n])
            if st:
                self._m_views[n].copy_(st['exp_avg'])
                self._v_views[n].copy_(st['exp_avg_sq'])
                self._step = max(self._step, int(st.get('step', self._step)))
        for n in flat.names:
            self.state.pop(flat.params[n], None)

    def zero_grad(self, set_to_none=True):
        """Never zero the flat gradient buffer in place: DistillationBox.forward has already written
        this step's gradients into it when the reference loop calls zero_grad() (mimic_runner.py:51).
        Views of the flat buffer are simply dropped; other gradients behave as in torch."""
        if self.flat is None:
            return super().zero_grad(set_to_none=set_to_none)
        own = {self.flat.grads[n].data_ptr() for n in self.flat.names}
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.data_ptr() in own or set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        self._step += 1
        group = self.param_groups[0]
        lr, (b1, b2), eps, wd = group['lr'], group['betas'], group['eps'], group['weight_decay']
        if self.flat is not None:
            f = self.flat
            for n in f.names:  # gradients produced outside the fused path (plain autograd) are gathered
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
                st['step'] = self._step
                g = p.grad.contiguous()
                ops.adam_step(p.data, g, st['exp_avg'], st['exp_avg_sq'], group['lr'], group['betas'][0],
                              group['betas'][1], group['eps'], group['weight_decay'], self.grad_scale,
                              self._step)
        return None

    # ---- checkpoints (torch.optim.Adam format) ---------------------------------------------------
    def _param_index(self):
        idx, i = {}, 0
        for group in self.param_groups:
            for p in group['params']:
                idx[id(p)] = i
                i += 1
        return idx

    def state_dict(self):
        sd = super().state_dict()
        if self.flat is not None and self._step > 0:
            idx = self._param_index()
            state = dict(sd['state'])
            for n in self.flat.names:
                p = self.flat.params[n]
                if id(p) in idx:
                    state[idx[id(p)]] = {'step': torch.tensor(float(self._step)),
                                         'exp_avg': self._m_views[n].detach().clone(),
                                         'exp_avg_sq': self._v_views[n].detach().clone()}
            sd['state'] = state
        sd['fused_step'] = self._step
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused_step = state_dict.pop('fused_step', None)
        super().load_state_dict(state_dict)
        steps = [int(st['step']) for st in self.state.values() if 'step' in st]
        self._step = int(fused_step) if fused_step is not None else (max(steps) if steps else 0)
        if self.flat is not None:  # move the loaded per-tensor moments into the flat buffers
            flat, self.flat = self.flat, None
            step = self._step
            self.attach(flat)
            self._step = step
