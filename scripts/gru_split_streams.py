"""Experiment: the 64 rows of a recurrent launch as TWO concurrent launches of 32 rows each on two streams, 16 units per
CTA (64 CTAs each, 128 SMs together): every CTA lands half the bytes per timestep and each handshake spans 64 CTAs.
python scripts/gru_split_streams.py [--steps 2000]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=2000)
ap.add_argument('--hidden', type=int, default=1024)
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()
t, h = a.steps, a.hidden
bf = torch.bfloat16
B = 64
w = (torch.randn(3 * h, h, device='cuda') / math.sqrt(h)).to(bf)
wt = w.t().contiguous()
b_hh = torch.zeros(3 * h, device='cuda')


def bufs(b):
    return dict(gi=torch.randn(b * t, 3 * h, device='cuda').to(bf), h_ext=torch.zeros(t + 1, b, h, dtype=bf, device='cuda'),
                hall=torch.zeros(b * t, h, dtype=bf, device='cuda'), gates=torch.empty(b * t, 4 * h, dtype=bf, device='cuda'),
                dh_out=(torch.randn(b * t, h, device='cuda') * 0.1).to(bf), dgi=torch.empty(b * t, 3 * h, dtype=bf, device='cuda'),
                dgh=torch.empty(b * t, 3 * h, dtype=bf, device='cuda'), dh0=torch.empty(b, h, device='cuda'), b=b)


def run(kind, d):
    b = d['b']
    if kind == 'fwd':
        ops.gru_forward(d['gi'], w, b_hh, d['h_ext'], d['hall'], torch.zeros(b, h, device='cuda'), d['gates'], b, t, h)
    else:
        ops.gru_backward(wt, d['h_ext'], d['gates'], d['dh_out'], d['dgi'], d['dgh'], d['dh0'], b, t, h)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


full = bufs(B)
halves = [bufs(B // 2), bufs(B // 2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def two_streams(kind):
    cur = torch.cuda.current_stream()
    for s, d in zip(streams, halves):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            run(kind, d)
    for s in streams:
        cur.wait_stream(s)


for kind in ('fwd', 'bwd'):
    ops.gru_units_per_cta = 8
    base = timed(lambda: run(kind, full))
    ops.gru_units_per_cta = 16
    one_half = timed(lambda: run(kind, halves[0]))
    both = timed(lambda: two_streams(kind))
    full16 = timed(lambda: run(kind, full))
    ops.gru_units_per_cta = 8
    half8 = timed(lambda: run(kind, halves[0]))
    print(f'{kind}: 64 rows, 8 units/CTA (128 CTAs): {1e3 * base / t:.2f} us/step | 64 rows, 16 units/CTA (64 CTAs): {1e3 * full16 / t:.2f} | '
          f'32 rows, 16 units/CTA alone: {1e3 * one_half / t:.2f} | 2 x 32 rows on two streams, 16 units/CTA: {1e3 * both / t:.2f} | '
          f'32 rows, 8 units/CTA alone: {1e3 * half8 / t:.2f}')
