"""Data-parallel training step over utterance slots.

One process per GPU.  Rank g owns batch slots [g*B/G, (g+1)*B/G) for the whole run, so the carried
hidden state of a slot never moves.  Per step the only exchange is
  * an all-reduce(sum) of the fp32 gradients, issued per *bucket* (one bucket per top-level module,
    in the order autograd finishes them: sample layer, tier 0, ..., top tier, mixer) from
    post-accumulate hooks so it overlaps the rest of the backward pass, and
  * a 2-scalar all-reduce (sum of NLL, number of valid rows) so that loss and gradient
    normalisation equal the single-process mean over all valid rows (model.py:283-284,
    runner.py:52) even when ranks hold different numbers of empty (reset == 2) slots.
Clipping (optimizer.py:12) is applied after averaging, inside the fused AdamClipped kernel.

Parameters and gradients live in two flat fp32 buffers (views are handed back to the modules), so
the optimizer is one kernel launch and each bucket is one contiguous NCCL call.
"""
import os
import weakref
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_slots(batch: int, world: int, rank: int):
    """Contiguous, sticky slot ownership: returns (first, last+1)."""
    if batch % world != 0:
        raise ValueError(f'global batch {batch} must be divisible by world size {world}')
    per = batch // world
    return rank * per, (rank + 1) * per


class FlatBuffers:
    """Re-homes every parameter (and its gradient) of ``model`` into flat fp32 buffers."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters()]
        self.names = [n for n, _ in model.named_parameters()]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + (s + 3) // 4 * 4)          # 16-byte aligned views
        total = self.offsets[-1]
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            view = self.flat_param[off: off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off: off + p.numel()].view_as(p)
        # buckets: contiguous ranges of parameters that share a top-level module prefix
        self.buckets = []                                                      # (name, start, end, param indices)
        cur, start, idxs = None, 0, []
        for i, n in enumerate(self.names):
            parts = n.split('.')
            key = '.'.join(parts[:2]) if parts[0] == 'frames_layers' else parts[0]
            if key != cur:
                if cur is not None:
                    self.buckets.append((cur, start, self.offsets[i], idxs))
                cur, start, idxs = key, self.offsets[i], []
            idxs.append(i)
        self.buckets.append((cur, start, self.offsets[-1], idxs))

    def zero_grad(self):
        self.flat_grad.zero_()


class FlatAdamClipped(torch.optim.Optimizer):
    """``AdamClipped`` (optimizer.py:6-14) over the flat buffers: clamp + Adam for ALL parameters in one launch.

    It is a real ``torch.optim.Optimizer``: ``param_groups[0]['lr']`` is read at every step, so
    ``ReduceLROnPlateau(trainer.optimizer)`` (runner.py:34-39,58-59) works, and ``state_dict()`` /
    ``load_state_dict()`` use torch.optim.Adam's layout (per-parameter ``step``, ``exp_avg``, ``exp_avg_sq``), so a
    checkpoint written by the reference's ``AdamClipped`` resumes here and vice versa.  The per-parameter moments are
    views of two flat buffers."""

    def __init__(self, flat: 'FlatBuffers', lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(flat.params, defaults)
        self.flat = flat
        self.exp_avg = torch.zeros_like(flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(flat.flat_param)
        self.steps = 0
        self._bind_state()

    def _bind_state(self):
        self._step_t = torch.tensor(float(self.steps))            # one shared CPU scalar: a single increment per step
        for p, off in zip(self.flat.params, self.flat.offsets):
            n = p.numel()
            self.state[p] = dict(step=self._step_t, exp_avg=self.exp_avg[off: off + n].view_as(p),
                                 exp_avg_sq=self.exp_avg_sq[off: off + n].view_as(p))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        from . import ops
        group = self.param_groups[0]
        if group.get('amsgrad') or group.get('weight_decay') or group.get('maximize'):
            raise RuntimeError('AdamClipped: amsgrad / weight_decay / maximize are not part of the reference path')
        self.steps += 1
        self._step_t += 1
        ops.adam_clipped(self.flat.flat_param, self.flat.flat_grad, self.exp_avg, self.exp_avg_sq, float(group['lr']),
                         group['betas'][0], group['betas'][1], group['eps'], self.steps, grad_scale=grad_scale)

    def state_dict(self):
        sd = super().state_dict()
        for st in sd['state'].values():                           # un-share the step scalar in the saved copy
            st['step'] = st['step'].clone()
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                       # torch re-creates the state tensors: copy them back
        steps = 0
        for p, off in zip(self.flat.params, self.flat.offsets):
            st = self.state.get(p)
            if not st:
                continue
            n = p.numel()
            self.exp_avg[off: off + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[off: off + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps = max(steps, int(st['step']))
        self.steps = steps
        self._bind_state()


class DataParallelTrainer:
    """forward + NLL + backward + bucketed all-reduce + fused AdamClipped for one rank.

    ``group=None`` with an uninitialised process group runs single-process (no collectives).
    Works with any backend; tests drive it with gloo on CPU tensors through ``reduce_only``."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, group=None, defer_buckets=None):
        self.model = model
        self.flat = FlatBuffers(model)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.optimizer = FlatAdamClipped(self.flat, lr=lr, betas=betas, eps=eps)
        self._pending: List = []
        self._ready = None
        # A bucket whose gradients are complete is not all-reduced at once but right AFTER the next recurrent backward launch
        # has been enqueued (torch orders the NCCL stream behind the work enqueued so far, so the all-reduce then runs after
        # that recurrence, next to the GEMMs that follow it).  Launched immediately it would still be resident - spinning for
        # the slowest rank - when the next persistent recurrent kernel is launched, and a cooperative launch waits for it:
        # 2 ms per step at 8 GPUs (profiles/r02_dp_trace_n8.txt).  Buckets that complete after the last recurrent launch
        # go out at the end of backward as before.
        if defer_buckets is None:
            defer_buckets = os.environ.get('SRNN_DP_DEFER', '1') != '0'
        self.defer_buckets = defer_buckets
        self._deferred: List[int] = []
        self._active = False
        self._install_hooks()

    def _install_hooks(self):
        if self.world == 1:
            return
        self._remaining = {}
        owner = {}
        for bi, (_, _, _, idxs) in enumerate(self.flat.buckets):
            for i in idxs:
                owner[i] = bi
        for i, p in enumerate(self.flat.params):
            p.register_post_accumulate_grad_hook(lambda _p, bi=owner[i]: self._param_ready(bi))
        from . import ops
        ref = weakref.WeakMethod(self._after_recurrent_launch)       # a dropped trainer must not keep receiving calls

        def listener():
            fn = ref()
            if fn is not None:
                fn()
        ops.rnn_backward_listeners.append(listener)

    def _param_ready(self, bi):
        self._remaining[bi] -= 1
        if self._remaining[bi] == 0:
            if self.defer_buckets:
                self._deferred.append(bi)
            else:
                self._launch_bucket(bi)

    def _after_recurrent_launch(self):
        """ops.rnn_backward_listeners callback: a recurrent backward kernel has just been enqueued."""
        if self._active:
            self._flush_deferred()

    def _flush_deferred(self):
        for bi in self._deferred:
            self._launch_bucket(bi)
        self._deferred = []

    def _launch_bucket(self, bi):
        _, start, end, _ = self.flat.buckets[bi]
        self._pending.append(dist.all_reduce(self.flat.flat_grad[start:end], op=dist.ReduceOp.SUM, group=self.group,
                                             async_op=True))
        self._launched.add(bi)

    def _begin(self):
        self._pending = []
        self._launched = set()
        self._deferred = []
        self._active = True
        if self.world > 1:
            self._remaining = {bi: len(b[3]) for bi, b in enumerate(self.flat.buckets)}

    def _finish_reduce(self):
        self._active = False
        if self.world == 1:
            return
        self._flush_deferred()
        for bi in range(len(self.flat.buckets)):                               # parameters that got no gradient
            if bi not in self._launched:
                self._launch_bucket(bi)
        for w in self._pending:
            w.wait()

    def reduce_only(self, stats: torch.Tensor):
        """All-reduce whatever sits in the flat gradient buffer plus the (sum_nll, n_valid) pair."""
        self._begin()
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        self._finish_reduce()
        return stats

    def step(self, x, y, utt_conds, info, reset, global_count=None):
        """One training step on this rank's slots.  Returns (global mean NLL, global valid rows).

        ``global_count``: the number of valid target rows over ALL ranks when the caller already knows
        it (e.g. no empty slots); the step then needs no device->host read and the returned loss is a
        device tensor.  Otherwise the (sum, count) pair is all-reduced and read back."""
        self._begin()
        self.flat.zero_grad()
        y_hat, tgt = self.model(x, y, utt_conds, info, reset)
        # local SUM of the NLL; the mean's 1/N_global is folded into the optimizer's grad_scale
        if y_hat.size(2) == 1:      # fused-loss mode: y_hat already holds log p(target)
            from .functional import NegSumFn
            local_sum = NegSumFn.apply(y_hat)
        else:
            local_sum = torch.nn.functional.nll_loss(y_hat.view(-1, y_hat.size(2)), tgt.view(-1), reduction='sum')
        stats = torch.stack([local_sum.detach(), torch.tensor(float(tgt.numel()), device=local_sum.device)])
        local_sum.backward()
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        self._finish_reduce()
        if global_count is None:
            total, count = stats.tolist()
            loss = total / max(count, 1.0)
        else:
            count = float(global_count)
            loss = stats[0] / count
        self.optimizer.step(grad_scale=1.0 / max(count, 1.0))
        return loss, int(count)

    # optimizer state in torch.optim.Adam layout (checkpoint interop with the reference's AdamClipped)
    @property
    def lr(self):
        return self.optimizer.param_groups[0]['lr']

    @property
    def steps(self):
        return self.optimizer.steps

    @property
    def exp_avg(self):
        return self.optimizer.exp_avg

    @property
    def exp_avg_sq(self):
        return self.optimizer.exp_avg_sq

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict)
