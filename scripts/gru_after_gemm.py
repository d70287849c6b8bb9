"""Does a burst of tensor-core GEMMs slow the latency-bound recurrent kernel that follows it?  (It does: the SM clock stays
reduced under the power cap for tens of milliseconds.)  python scripts/gru_after_gemm.py"""
import math, os, sys, torch
sys.path.insert(0, os.getcwd())
from samplernn_pase_b200 import ops
b, t, h = 64, 4000, 1024
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
ops.gru_forward(gi, w, b_hh, h_ext, hall, torch.zeros(b, h, device='cuda'), gates, b, t, h)
m = 1024000
cat = torch.randn(m, 2 * h, device='cuda').to(bf)
wc = (torch.randn(h, 2 * h, device='cuda') * 0.02).to(bf)
h1 = torch.empty(m, h, dtype=bf, device='cuda')


def bwd():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gru_backward(wt, h_ext, gates, dh_out, dgi, dgh, dh0, b, t, h); e1.record()
    return e0, e1


bwd(); torch.cuda.synchronize()
for burst in (0, 4, 10, 20):
    for _ in range(burst):
        ops.gemm_nt(cat, wc, h1, m, h, 2 * h, 2 * h, 2 * h, h, relu=True)
    e0, e1 = bwd()
    torch.cuda.synchronize()
    print(f'bwd T=4000 after {burst} x 3.4 ms GEMMs: {e0.elapsed_time(e1):.2f} ms = {1e3 * e0.elapsed_time(e1) / t:.2f} us/step')
    torch.cuda.synchronize()
