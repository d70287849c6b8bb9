"""Kernel timeline of the data-parallel training step (CUPTI through torch.profiler): where the NCCL all-reduce kernels run
relative to the recurrent kernels, and how long each recurrent launch waited for its SMs.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_trace.py
Config 2, 64 slots per rank; 3 warm-up steps, then 2 steps under the profiler (every rank profiles; ranks 0 and N-1 print).
Numbers under the profiler are not bench values; the point is the ORDER and OVERLAP of kernels."""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                              # noqa: E402
from samplernn_pase_b200 import SampleRNNModel, synthetic                 # noqa: E402
from samplernn_pase_b200.parallel import DataParallelTrainer              # noqa: E402

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(1234)
model = SampleRNNModel(fused_loss=True, **bench.model_kwargs()).to(dev)
trainer = DataParallelTrainer(model, lr=1e-4)
fs, rf, b, seq = int(model.frame_size), int(model.receptive_field), 64, bench.seq_len_default()
wav, conds, spk = synthetic.synthetic_utterances(fs, rf, seq, b, 2, seed=4321 + rank)
info = [{'speaker': {'index': int(s)}} for s in spk]
chunks = [tuple(t.to(dev) for t in synthetic.chunk_of(fs, rf, seq, wav, conds, k)) for k in range(2)]
resets = [torch.ones(b, dtype=torch.int64), torch.zeros(b, dtype=torch.int64)]


def step(s):
    x, y, c = chunks[s % 2]
    trainer.step(x, y, c, info, resets[s % 2], global_count=b * rf * world)


for s in range(3):
    step(s)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(3, 5):
        step(s)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
lines = []


def kind(name):
    if 'nccl' in name.lower():
        return 'nccl'
    if 'gru_kernel' in name:
        return 'gru'
    return 'other'


last_end = None
last_main_end = None                                   # end of the last non-NCCL kernel (the training stream)
nccl = [e for e in ev if kind(e.name) == 'nccl']
for e in ev:
    k = kind(e.name)
    if k == 'gru':
        gap = (e.time_range.start - last_end) if last_end is not None else 0
        gap_main = (e.time_range.start - last_main_end) if last_main_end is not None else 0
        over = [n for n in nccl if n.time_range.start < e.time_range.end and n.time_range.end > e.time_range.start]
        before = [n for n in nccl if n.time_range.end <= e.time_range.start and e.time_range.start - n.time_range.end < 200]
        short = e.name.split('<')[1].split('>')[0] if '<' in e.name else e.name
        lines.append(f'  {(e.time_range.start - t0) / 1e3:9.3f} ms  gru_kernel<{short}>  {(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms, '
                     f'started {gap_main:7.1f} us after the previous kernel of the training stream ended ({gap:.1f} us after ANY kernel); NCCL kernels running during it: {len(over)}'
                     + (f' (first overlaps from {(over[0].time_range.start - e.time_range.start) / 1e3:+.3f} ms for '
                        f'{(over[0].time_range.end - over[0].time_range.start) / 1e3:.3f} ms)' if over else '')
                     + (f'; an NCCL kernel ended {e.time_range.start - before[-1].time_range.end:.0f} us before it' if before else ''))
    elif k == 'nccl':
        lines.append(f'  {(e.time_range.start - t0) / 1e3:9.3f} ms  {e.name[:48]:48s} {(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms')
    last_end = max(last_end, e.time_range.end) if last_end is not None else e.time_range.end
    if k != 'nccl':
        last_main_end = max(last_main_end, e.time_range.end) if last_main_end is not None else e.time_range.end
span = (ev[-1].time_range.end - t0) / 1e3
busy = sum(e.time_range.end - e.time_range.start for e in ev if kind(e.name) != 'nccl') / 1e3
out = [f'rank {rank} of {world}: 2 profiled steps span {span:.2f} ms; sum of non-NCCL kernel time {busy:.2f} ms; '
       f'{len(nccl)} NCCL kernels, {sum(n.time_range.end - n.time_range.start for n in nccl) / 1e3:.2f} ms in total'] + lines
if rank in (0, world - 1):
    os.makedirs('gpurun_out', exist_ok=True)
    with open(f'gpurun_out/dp_trace_n{world}_rank{rank}.txt', 'w') as f:
        f.write('\n'.join(out) + '\n')
    if rank == 0:
        print('\n'.join(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
