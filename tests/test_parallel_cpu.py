"""World-size-2 gloo checks of the data-parallel host logic (no GPU): slot sharding, flat
parameter/gradient buffers, per-module buckets and the bucketed all-reduce + (sum, count) exchange."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from samplernn_pase_b200.parallel import DataParallelTrainer, FlatBuffers, shard_slots


class Tiny(torch.nn.Module):
    """Same top-level module names as SampleRNNModel so the bucket rule is exercised."""

    def __init__(self):
        super().__init__()
        self.conds_mixer = torch.nn.Linear(5, 3)
        self.frames_layers = torch.nn.ModuleList([torch.nn.Linear(3, 3), torch.nn.Linear(3, 3)])
        self.sample_layer = torch.nn.Linear(3, 7)

    def forward(self, x):
        x = self.conds_mixer(x)
        for layer in self.frames_layers:
            x = torch.tanh(layer(x))
        return self.sample_layer(x)


def test_shard_slots():
    assert shard_slots(128, 8, 3) == (48, 64)
    assert [shard_slots(64, 4, r) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    with pytest.raises(ValueError):
        shard_slots(10, 4, 0)


def test_flat_buffers_views_and_buckets():
    torch.manual_seed(0)
    m = Tiny()
    before = [p.detach().clone() for p in m.parameters()]
    flat = FlatBuffers(m)
    assert [b[0] for b in flat.buckets] == ['conds_mixer', 'frames_layers.0', 'frames_layers.1', 'sample_layer']
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, m.parameters()))
    assert all(p.data.data_ptr() >= flat.flat_param.data_ptr() for p in m.parameters())
    assert all(off % 4 == 0 for off in flat.offsets)                     # 16-byte aligned views
    m(torch.randn(4, 5)).sum().backward()                                  # autograd accumulates INTO the flat buffer
    assert float(flat.flat_grad.abs().sum()) > 0
    for p, off in zip(flat.params, flat.offsets):
        assert torch.equal(flat.flat_grad[off: off + p.numel()].view_as(p), p.grad)
    flat.zero_grad()
    assert float(flat.flat_grad.abs().sum()) == 0 and all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


def _worker(rank, world, port, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    m = Tiny()
    trainer = DataParallelTrainer(m)
    # every rank works on its own slots: rank r holds rows [lo, hi) of a global batch of 8
    gen = torch.Generator().manual_seed(1)
    x_all = torch.randn(8, 5, generator=gen)
    lo, hi = shard_slots(8, world, rank)
    trainer._begin()
    trainer.flat.zero_grad()
    loss = m(x_all[lo:hi]).pow(2).sum()
    loss.backward()                                                        # hooks launch the bucket all-reduces
    stats = torch.tensor([float(loss), float(hi - lo)])
    dist.all_reduce(stats)
    trainer._finish_reduce()
    results[rank] = (trainer.flat.flat_grad.clone(), stats.clone(), sorted(trainer._launched))
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    torch.manual_seed(0)
    m = Tiny()
    flat = FlatBuffers(m)
    gen = torch.Generator().manual_seed(1)
    x_all = torch.randn(8, 5, generator=gen)
    loss = m(x_all).pow(2).sum()
    loss.backward()
    for rank in range(2):
        grad, stats, launched = results[rank]
        assert torch.allclose(grad, flat.flat_grad, atol=1e-5)             # sum over ranks == single-process gradient
        assert abs(float(stats[0]) - float(loss)) < 1e-4 and int(stats[1]) == 8
        assert launched == [0, 1, 2, 3]                                    # one all-reduce per bucket


def _worker_empty_shard(rank, world, port, results):
    """Rank 1 holds only empty slots (reset == 2 everywhere: the tail of every epoch, slot ownership is sticky): its
    filtered log-probability tensor has zero rows.  The loss path must still produce a graph-connected zero so that
    backward runs and every bucket all-reduce is matched on both ranks (no hang, no exception)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from samplernn_pase_b200.functional import NegSumFn
    torch.manual_seed(0)
    m = Tiny()
    trainer = DataParallelTrainer(m)
    gen = torch.Generator().manual_seed(1)
    x_all = torch.randn(8, 5, generator=gen)
    lo, hi = shard_slots(8, world, rank)
    keep = torch.arange(hi - lo) if rank == 0 else torch.zeros(0, dtype=torch.int64)
    trainer._begin()
    trainer.flat.zero_grad()
    logp_t = m(x_all[lo:hi])[:, :1].index_select(0, keep).unsqueeze(2)        # (B_valid, 1, 1) like fused-loss forward
    if logp_t.numel():
        local_sum = -logp_t.sum()          # (the CUDA reduction kernel is not available on CPU; same value)
    else:
        local_sum = NegSumFn.apply(logp_t)                                     # the path under test
    assert local_sum.requires_grad
    local_sum.backward()
    stats = torch.tensor([float(local_sum), float(logp_t.numel())])
    dist.all_reduce(stats)
    trainer._finish_reduce()
    results[rank] = (trainer.flat.flat_grad.clone(), stats.clone(), sorted(trainer._launched))
    dist.destroy_process_group()


def test_rank_with_only_empty_slots_keeps_collectives_matched():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_empty_shard, args=(2, port, results), nprocs=2, join=True)
    torch.manual_seed(0)
    m = Tiny()
    flat = FlatBuffers(m)
    gen = torch.Generator().manual_seed(1)
    x_all = torch.randn(8, 5, generator=gen)
    loss = -m(x_all[:4])[:, :1].sum()                                          # only rank 0's slots are valid
    loss.backward()
    for rank in range(2):
        grad, stats, launched = results[rank]
        assert torch.allclose(grad, flat.flat_grad, atol=1e-5)
        assert abs(float(stats[0]) - float(loss)) < 1e-4 and int(stats[1]) == 4
        assert launched == [0, 1, 2, 3]


def test_flat_adam_clipped_state_dict_uses_torch_adam_layout():
    """The trainer's optimizer is a torch Optimizer: lr is read from param_groups (ReduceLROnPlateau works on it,
    runner.py:34-39) and its state_dict has torch.optim.Adam's layout, loadable by / from a plain Adam."""
    from samplernn_pase_b200.parallel import FlatAdamClipped
    torch.manual_seed(0)
    m = Tiny()
    flat = FlatBuffers(m)
    opt = FlatAdamClipped(flat, lr=1e-3)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.5, patience=0)
    sched.step(1.0); sched.step(2.0)
    assert abs(opt.param_groups[0]['lr'] - 5e-4) < 1e-12
    ref = torch.optim.Adam(Tiny().parameters(), lr=1e-3)
    sd = opt.state_dict()
    assert set(sd.keys()) == set(ref.state_dict().keys())
    assert set(sd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    # round trip through a plain Adam state (moments as separate tensors) back into the flat buffers
    for i, st in sd['state'].items():
        st['exp_avg'] = torch.full_like(st['exp_avg'], float(i + 1))
        st['exp_avg_sq'] = torch.full_like(st['exp_avg_sq'], 0.5)
        st['step'] = torch.tensor(7.0)
    opt.load_state_dict(sd)
    assert opt.steps == 7
    for i, (p, off) in enumerate(zip(flat.params, flat.offsets)):
        assert bool((opt.exp_avg[off: off + p.numel()] == i + 1).all())
        assert opt.state[p]['exp_avg'].data_ptr() == opt.exp_avg[off: off + p.numel()].data_ptr()


def _worker_deferred(rank, world, port, results):
    """Deferred buckets: a completed bucket is held back until a recurrent backward launch has been enqueued
    (``ops.rnn_backward_listeners``) or backward ends; with ``defer_buckets=False`` it goes out from the hook."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from samplernn_pase_b200 import ops
    torch.manual_seed(0)
    m = Tiny()
    trainer = DataParallelTrainer(m, defer_buckets=True)
    x = torch.randn(4, 5, generator=torch.Generator().manual_seed(rank))
    seen = {}
    # a "recurrent backward launch" in the middle of backward: when the gradient of frames_layers[1]'s input is computed
    def mid(_grad):
        seen['before'] = (sorted(trainer._launched), list(trainer._deferred))
        for fn in ops.rnn_backward_listeners:
            fn()
        seen['after'] = (sorted(trainer._launched), list(trainer._deferred))
    trainer._begin()
    trainer.flat.zero_grad()
    h = torch.tanh(m.frames_layers[0](m.conds_mixer(x)))
    h.register_hook(mid)
    m.sample_layer(torch.tanh(m.frames_layers[1](h))).pow(2).sum().backward()
    end = (sorted(trainer._launched), list(trainer._deferred))
    trainer._finish_reduce()
    deferred_grad = trainer.flat.flat_grad.clone()
    # the same step with immediate launches gives the same reduced gradient
    torch.manual_seed(0)
    m2 = Tiny()
    t2 = DataParallelTrainer(m2, defer_buckets=False)
    t2._begin()
    t2.flat.zero_grad()
    m2(x).pow(2).sum().backward()
    immediate = sorted(t2._launched)
    t2._finish_reduce()
    results[rank] = (seen, end, sorted(trainer._launched), immediate, bool(torch.allclose(deferred_grad, t2.flat.flat_grad, atol=1e-6)))
    dist.destroy_process_group()


def test_deferred_buckets_go_out_behind_a_recurrent_launch():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_deferred, args=(2, port, results), nprocs=2, join=True)
    for rank in range(2):
        seen, end, launched, immediate, same = results[rank]
        # buckets in FlatBuffers order: conds_mixer 0, frames_layers.0 1, frames_layers.1 2, sample_layer 3
        assert seen['before'] == ([], [3, 2])            # sample layer and the upper tier are complete but held back
        assert seen['after'] == ([2, 3], [])             # the "recurrent launch" releases them
        assert end == ([2, 3], [1, 0])                   # the rest waits for the end of backward
        assert launched == [0, 1, 2, 3] and immediate == [0, 1, 2, 3] and same
