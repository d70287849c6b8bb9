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
