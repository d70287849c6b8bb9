"""The adapt data-gradient GEMM as the training step runs it (m x 1024 x 256, ReLU bit-mask gate, fused column sums) and the
plain variants, for ncu / timing.  python scripts/ncu_k256.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops   # noqa: E402

bf = torch.bfloat16
m, h, q = int(os.environ.get('ROWS', 1024000)), 1024, 256
dlog = (torch.randn(m, q, device='cuda') * 0.01).to(bf)
w3t = (torch.randn(h, q, device='cuda') * 0.05).to(bf)
dh2 = torch.empty(m, h, dtype=bf, device='cuda')
mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (m, h // 32), dtype=torch.int32, device='cuda')
cs = torch.zeros(h, device='cuda')


def timed(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f'{name:60s} {best:7.3f} ms   {2.0 * m * h * q / best / 1e9:7.1f} TFLOP/s   {(m * q * 2 + m * h * 2) / best / 1e6:7.1f} GB/s (A + C)')


timed('gate mask + colsum (as in the step)', lambda: ops.gemm_nt(dlog, w3t, dh2, m, h, q, q, q, h, gate_mask=mask, colsum=cs))
timed('gate mask, no colsum', lambda: ops.gemm_nt(dlog, w3t, dh2, m, h, q, q, q, h, gate_mask=mask))
timed('plain store', lambda: ops.gemm_nt(dlog, w3t, dh2, m, h, q, q, q, h))
timed('colsum only', lambda: ops.gemm_nt(dlog, w3t, dh2, m, h, q, q, q, h, colsum=cs))
