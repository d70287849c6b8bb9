"""Timing decomposition of the persistent GRU kernels (CUDA events, us per timestep).
python scripts/gru_microbench.py [--batch 64] [--steps 1000] [--hidden 1024] [--flags 0,1,2,3,4,7]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=1000)
ap.add_argument('--hidden', type=int, default=1024)
ap.add_argument('--flags', default='0,16,32')   # 0 = relaxed publish + validated read, 16 = strict protocol, 32 = per-K-block landing, bits 8.. force a cluster size
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--ts-flags', default='0,16')              # timeline(s) of CTA 0 for these flag values
ap.add_argument('--units', type=int, default=0)          # units per CTA (0 = default 8)
a = ap.parse_args()
b, t, h = a.batch, a.steps, a.hidden
ops.gru_units_per_cta = a.units
bf = torch.bfloat16
gi = torch.randn(b * t, 3 * h, device='cuda').to(bf)
w = (torch.randn(3 * h, h, device='cuda') / math.sqrt(h)).to(bf)
wt = w.t().contiguous()
b_hh = torch.zeros(3 * h, device='cuda')
h_ext = torch.zeros(t + 1, b, h, dtype=bf, device='cuda')
hall = torch.zeros(b * t, h, dtype=bf, device='cuda')
gates = torch.empty(b * t, 4 * h, dtype=bf, device='cuda')
dh_out = (torch.randn(b * t, h, device='cuda') * 0.1).to(bf)
dgi = torch.empty(b * t, 3 * h, dtype=bf, device='cuda')
dgh = torch.empty(b * t, 3 * h, dtype=bf, device='cuda')
dh0 = torch.empty(b, h, device='cuda')


def run(kind):
    if kind == 'fwd':
        ops.gru_forward(gi, w, b_hh, h_ext, hall, torch.zeros(b, h, device='cuda'), gates, b, t, h)
    else:
        ops.gru_backward(wt, h_ext, gates, dh_out, dgi, dgh, dh0, b, t, h)


for flags in [int(f) for f in a.flags.split(',')]:
    ops.gru_tuning_flags = flags
    for kind in ('fwd', 'bwd'):
        try:
            run(kind)
        except RuntimeError as e:
            print(f'flags={flags} {kind}: {str(e)[-80:]}')
            continue
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(kind); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f'flags={flags} {kind}: {best:8.3f} ms  {1e3 * best / t:6.2f} us/step   (B={b} T={t} H={h})  repeated attempts in the last launch: {int(ops.gru_last_sync[32])}')
# pipeline timestamps of CTA 0 (forward): cycles relative to the end of the grid wait
names = ['wait_done', 'tma_issued', 'last_landed', 'mma_commit', 'epi_acc_full', 'pushed/1st_land', 'epi_part_rdy', 'epi_publish']
order = [0, 1, 2, 3, 4, 7, 5, 6]
for flags in [int(f) for f in a.ts_flags.split(',') if f]:
    for kind in ('fwd', 'bwd'):
        ts = torch.zeros(256, 8, dtype=torch.int64, device='cuda')
        ops.gru_tuning_flags = flags
        ops.gru_debug_ts = ts
        run(kind)
        torch.cuda.synchronize()
        ops.gru_debug_ts = None
        t_ = ts.cpu()
        print(f'timeline flags={flags} {kind}')
        print('step  ' + '  '.join(f'{n:>12s}' for n in names) + '   next_wait_done')
        for s_ in range(20, 26):
            base = int(t_[s_, 0])
            print(f'{s_:4d}  ' + '  '.join(f'{int(t_[s_, i]) - base:12d}' for i in order) + f'   {int(t_[s_ + 1, 0]) - base:10d}')
ops.gru_tuning_flags = 0
# skew between CTAs: global timer (ns) of every CTA at timestep 24 of the forward and of the backward kernel
for kind in ('fwd', 'bwd'):
    ts = torch.zeros(256, 8, dtype=torch.int64, device='cuda')
    ops.gru_tuning_flags = 128
    ops.gru_debug_ts = ts
    run(kind)
    torch.cuda.synchronize()
    ops.gru_debug_ts = None
    ops.gru_tuning_flags = 0
    t_ = ts.cpu()
    n_cta = int((t_[:, 0] > 0).sum())
    t0 = int(t_[:n_cta, 0].min())
    print(f'{kind}: per-CTA global timer at step 24 ({n_cta} CTAs), ns after the first CTA left the grid wait')
    for i, nm in zip(order, names):
        col = t_[:n_cta, i] - t0
        if int(t_[:n_cta, i].max()) <= 0:
            continue
        print(f'{nm:>14s}: min {int(col.min()):6d}  median {int(col.median()):6d}  max {int(col.max()):6d}')
