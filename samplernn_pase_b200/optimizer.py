"""Drop-in for ``samplernn_pase.optimizer.AdamClipped`` (optimizer.py:6-14): every gradient is
clamped to [-1, 1] and a plain Adam update follows.  Clamp + moment update + parameter update
run as ONE fused pass per tensor (``srnn_adam_clipped``) instead of a hardtanh launch per tensor
followed by torch's foreach Adam.  State layout (``step``, ``exp_avg``, ``exp_avg_sq``) and
``param_groups`` are torch.optim.Adam's, so optimizer checkpoints interoperate.
"""
import torch

from . import ops


class AdamClipped(torch.optim.Adam):

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None and closure is not Ellipsis:       # the reference's default is `...`
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group.get('amsgrad') or group.get('weight_decay') or group.get('maximize'):
                raise RuntimeError('AdamClipped: amsgrad / weight_decay / maximize are not part of the reference path')
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    # optimizer.py:12 would raise TypeError inside F.hardtanh(None); keep that contract
                    raise TypeError('AdamClipped.step: a parameter has no gradient (reference raises here too)')
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = torch.tensor(0.0)
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state['step'] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adam_clipped(p.data, g, state['exp_avg'], state['exp_avg_sq'], float(group['lr']), beta1, beta2,
                                 group['eps'], int(state['step']))
        return loss
