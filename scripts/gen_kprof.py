"""Kernel-level device times of batched generation via torch.profiler (CUPTI); diagnostic only.
python scripts/gen_kprof.py [--batch 256] [--frames 16] [--eager]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from samplernn_pase_b200 import SampleRNNModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--frames', type=int, default=16)
ap.add_argument('--eager', action='store_true')
ap.add_argument('--no-pdl', action='store_true')   # exclusive kernel durations (with PDL a kernel's time includes waiting for its predecessor)
a = ap.parse_args()
torch.manual_seed(0)
if a.no_pdl:
    from samplernn_pase_b200 import generate as _g
    _g._PDL = False
model = SampleRNNModel(**bench.model_kwargs(a.frames)).cuda()
utt = torch.randn(a.batch, a.frames, 43).cuda()
info = [{'speaker': {'index': i % 126}} for i in range(a.batch)]
model.test(utt[:, :3], info, use_graphs=not a.eager)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model.test(utt, info, use_graphs=not a.eager)
    torch.cuda.synchronize()
n = a.frames * int(model.frame_size)
rows = []
for ev in prof.key_averages():
    if ev.device_time_total > 0:
        rows.append((ev.device_time_total, ev.count, ev.key))
tot = sum(r[0] for r in rows)
print(f'{n} sample steps; total device time {tot / n:.1f} us per sample step')
for t, c, k in sorted(rows, reverse=True)[:40]:
    print(f'{t / n:8.2f} us/sample  x{c / n:6.2f}  {t / c:8.2f} us each  {k[:110]}')
