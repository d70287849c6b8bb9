"""Can the latency-bound recurrence share the GPU with throughput-bound GEMMs?  Runs the persistent GRU
kernels (8 or 16 units per CTA = 128 or 64 CTAs) on one stream and capped-grid TN GEMMs on another and
reports each side's time alone and together."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops   # noqa: E402

bf = torch.bfloat16
b, t, h = 64, 1000, 1024
m = 1024000
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
x1 = torch.randn(m, h, device='cuda').to(bf)
x2 = torch.randn(m, h, device='cuda').to(bf)
dw = torch.zeros(h, h, device='cuda')
side = torch.cuda.Stream()


def gru(kind):
    if kind == 'fwd':
        ops.gru_forward(gi, w, b_hh, h_ext, hall, torch.zeros(b, h, device='cuda'), gates, b, t, h)
    else:
        ops.gru_backward(wt, h_ext, gates, dh_out, dgi, dgh, dh0, b, t, h)


def gemms(n):
    for _ in range(n):
        ops.gemm_tn(x1, x2, dw, h, h, m, h, h, h)


def timed(fn, stream=None):
    s = stream or torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s); fn(); e1.record(s)
    return e0, e1


for units, cap in ((8, 20), (16, 84)):
    ops.gru_units_per_cta = units
    for kind in ('fwd', 'bwd'):
        gru(kind); torch.cuda.synchronize()
        a = timed(lambda: gru(kind)); torch.cuda.synchronize()
        alone_gru = a[0].elapsed_time(a[1])
        n_gemm = 3
        ops.gemm_max_ctas = cap
        with torch.cuda.stream(side):
            gemms(1)
        torch.cuda.synchronize()
        g = timed(lambda: gemms(n_gemm), side); torch.cuda.synchronize()
        alone_gemm = g[0].elapsed_time(g[1])
        torch.cuda.synchronize()
        g = timed(lambda: gemms(n_gemm), side)          # GEMMs first so they hold their SMs, then the GRU
        a = timed(lambda: gru(kind))
        torch.cuda.synchronize()
        ops.gemm_max_ctas = 0
        full = timed(lambda: gemms(n_gemm)); torch.cuda.synchronize()
        print(f'units/CTA {units:2d} ({h // units} CTAs) {kind}: GRU alone {alone_gru:7.2f} ms | {n_gemm} TN GEMMs on <= {cap} CTAs alone '
              f'{alone_gemm:7.2f} ms (uncapped {full[0].elapsed_time(full[1]):6.2f}) | together: GRU {a[0].elapsed_time(a[1]):7.2f} ms, '
              f'GEMMs {g[0].elapsed_time(g[1]):7.2f} ms')
ops.gru_units_per_cta = 8
