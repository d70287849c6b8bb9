"""Per-C-call device-time breakdown of one training step (CUDA events around every C-ABI call; no
profiler attached).  python scripts/profile_step.py [--seq-len 1000] [--batch 64] > gpurun_out/step_breakdown.txt"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                  # noqa: E402
from samplernn_pase_b200 import SampleRNNModel, _lib, ops, synthetic   # noqa: E402
from samplernn_pase_b200.parallel import DataParallelTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--seq-len', type=int, default=1000)
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=2)
a = ap.parse_args()
torch.manual_seed(1234)
model = SampleRNNModel(fused_loss=True, **bench.model_kwargs(a.seq_len)).cuda()
trainer = DataParallelTrainer(model)
fs, rf = int(model.frame_size), int(model.receptive_field)
wav, conds, spk = synthetic.synthetic_utterances(fs, rf, a.seq_len, a.batch, a.steps + 2)
info = [{'speaker': {'index': int(s)}} for s in spk]
rows = a.batch * rf
for k in range(a.steps + 2):
    x, y, c = (t.cuda() for t in synthetic.chunk_of(fs, rf, a.seq_len, wav, conds, k))
    if k == 2:
        torch.cuda.synchronize()
        _lib.profile_log = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
    trainer.step(x, y, c, info, torch.ones(a.batch, dtype=torch.int64) if k == 0 else torch.zeros(a.batch, dtype=torch.int64),
                 global_count=rows)
t1.record()
torch.cuda.synchronize()
log, _lib.profile_log = _lib.profile_log, None
total = t0.elapsed_time(t1) / a.steps
agg = collections.OrderedDict()
for name, note, s, e in log:
    key = (name, note if 'gemm' in name or 'gru' in name else '')
    d = agg.setdefault(key, [0, 0.0])
    d[0] += 1
    d[1] += s.elapsed_time(e)
print(f'step {total:.2f} ms  ({rows / total * 1e3 / 1e6:.2f} M samples/s), {len(log) // a.steps} C calls/step, '
      f'sum of calls {sum(v[1] for v in agg.values()) / a.steps:.2f} ms')
for (name, note), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{ms / a.steps:9.3f} ms  {100 * ms / a.steps / total:5.1f}%  x{n // a.steps:<3d} {name:24s} {note}')
