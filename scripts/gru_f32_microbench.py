"""fp32-mode recurrence: us per timestep (one launch per timestep, SIMT fp32).  python scripts/gru_f32_microbench.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops    # noqa: E402

b, t, h = 64, 400, 1024
gi = torch.randn(b * t, 3 * h, device='cuda')
w = torch.randn(3 * h, h, device='cuda') / math.sqrt(h)
bias = torch.zeros(3 * h, device='cuda')
h0 = torch.zeros(b, h, device='cuda')
dh = torch.randn(b * t, h, device='cuda') * 0.1
for rep in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    hall, gates = ops.gru_forward_f32(gi, w, bias, h0.clone(), b, t, h)
    e[1].record()
    ops.gru_backward_f32(w, gates, hall, h0, dh, b, t, h)
    e[2].record()
    torch.cuda.synchronize()
print(f'fp32 recurrence B={b} T={t} H={h}: forward {1e3 * e[0].elapsed_time(e[1]) / t:.1f} us per timestep, '
      f'backward {1e3 * e[1].elapsed_time(e[2]) / t:.1f}')
