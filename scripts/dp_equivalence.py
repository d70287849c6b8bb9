"""Data-parallel equivalence on real GPUs (SURVEY section 4 'Distributed: 1 vs N ranks -> same averaged gradients'):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_equivalence.py
Every rank trains its shard of a 32-slot global batch (config-2 model at full width, two chunks with carry) through
DataParallelTrainer; rank 0 also trains the whole batch alone on a second model and compares, after each step, the
all-reduced flat gradient and the updated parameters.  (The reduction orders differ - per-rank split-K atomics, NCCL sum -
so equality is to fp32/bf16 accumulation noise, not bitwise; all RANKS must be bitwise identical among themselves.)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                              # noqa: E402
from samplernn_pase_b200 import SampleRNNModel, synthetic                 # noqa: E402
from samplernn_pase_b200.parallel import DataParallelTrainer, shard_slots  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
seq, glob = 64, 32
torch.manual_seed(1234)
model = SampleRNNModel(fused_loss=True, **bench.model_kwargs(seq)).to(dev)
ref_model = None
if rank == 0:
    torch.manual_seed(1234)
    ref_model = SampleRNNModel(fused_loss=True, **bench.model_kwargs(seq)).to(dev)
trainer = DataParallelTrainer(model, lr=1e-4)
fs, rf = int(model.frame_size), int(model.receptive_field)
wav, conds, spk = synthetic.synthetic_utterances(fs, rf, seq, glob, 2)
lo, hi = shard_slots(glob, world, rank)
info = [{'speaker': {'index': int(s)}} for s in spk]
solo = None
if rank == 0:
    solo_group = None
    solo = DataParallelTrainer.__new__(DataParallelTrainer)               # single-process trainer on the second model
    from samplernn_pase_b200.parallel import FlatAdamClipped, FlatBuffers
    solo.model, solo.flat, solo.group, solo.world = ref_model, FlatBuffers(ref_model), None, 1
    solo.optimizer = FlatAdamClipped(solo.flat, lr=1e-4)
    solo._pending, solo._ready = [], None
for k in range(2):
    x, y, c = (t.to(dev) for t in synthetic.chunk_of(fs, rf, seq, wav, conds, k))
    reset_all = [1] * glob if k == 0 else [0] * glob
    if k == 1:
        reset_all[3] = 1                                                   # a mid-stream new utterance
        reset_all[glob - 2] = 2                                            # and an empty slot on the last rank
    reset = torch.tensor(reset_all[lo:hi])
    inf = [None if r == 2 else i for i, r in zip(info[lo:hi], reset_all[lo:hi])]
    loss, n = trainer.step(x[lo:hi], y[lo:hi], c[lo:hi], inf, reset)
    chk = trainer.flat.flat_param.view(torch.int32).to(torch.int64).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        inf_all = [None if r == 2 else i for i, r in zip(info, reset_all)]
        loss1, n1 = solo.step(x, y, c, inf_all, torch.tensor(reset_all))
        g, g1 = trainer.flat.flat_grad.double(), solo.flat.flat_grad.double()
        p, p1 = trainer.flat.flat_param.double(), solo.flat.flat_param.double()
        rel_g = float((g - g1).norm() / g1.norm())
        cos_g = float((g @ g1) / (g.norm() * g1.norm()))
        print(f'chunk {k}: world {world}: loss {float(loss):.6f} vs single-process {float(loss1):.6f}; valid rows {n} vs {n1}; '
              f'summed gradient rel-L2 {rel_g:.3e} cos {cos_g:.8f}; max |param diff| {float((p - p1).abs().max()):.3e}; '
              f'ranks bitwise identical: {all(int(a) == int(allc[0]) for a in allc)}')
        assert n == n1 and abs(float(loss) - float(loss1)) < 1e-4 * float(loss1)
        assert rel_g < 2e-2 and cos_g > 0.9998 and all(int(a) == int(allc[0]) for a in allc)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print('dp equivalence ok')
