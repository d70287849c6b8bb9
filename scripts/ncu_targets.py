"""Launches the hot kernels once each at BASELINE config-2 row counts so that `ncu --set full` can
capture them in isolation (the persistent cluster GRU kernels are excluded: ncu cannot replay a
cooperative cluster launch - it reports LaunchFailed).  Prints CUDA-event times as a plain run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops   # noqa: E402

bf = torch.bfloat16
m, h, q = int(os.environ.get('ROWS', 1024000)), 1024, 256
dev = 'cuda'


def timed(name, fn, flops=None, nbytes=None):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    extra = f'{flops / ms / 1e9:8.1f} TFLOP/s' if flops else f'{nbytes / ms / 1e6:8.1f} GB/s'
    print(f'{name:50s} {ms:8.3f} ms  {extra}')


cat = torch.randn(m, 2 * h, device=dev).to(bf)
w = (torch.randn(h, 2 * h, device=dev) * 0.02).to(bf)
bias = torch.zeros(h, device=dev)
h1 = torch.empty(m, h, dtype=bf, device=dev)
# comb_layer forward as the model runs it: A = [overlapping one-hot windows (K1 = 4*256) | upper (K2 = H)], weight [T' | W_u],
# frame-rate conditioning term + ReLU in the epilogue
nb = max(1, m // 16000)
rf = m // nb // 16 * 16
win = rf + 3
codes = torch.randint(0, 256, (nb, win), dtype=torch.uint8, device=dev)
onehot_c = ops.onehot_rows(codes)
upper_c = torch.randn(nb * rf, h, device=dev).to(bf)
kc = 4 * q + h
w_cat = (torch.randn(h, kc, device=dev) * 0.02).to(bf)
cterm = torch.randn(nb * (rf // 16), h, device=dev).to(bf)
h1c = torch.empty(nb * rf, h, dtype=bf, device=dev)
timed('NT comb fwd  m x 1024 x (1024 one-hot + 1024)', lambda: ops.gemm_nt(
    onehot_c, w_cat, h1c, rf, h, kc, q, kc, h, batch=nb, a_bs=win * q, c_bs=rf * h, aux=cterm, ldaux=h, aux_bs=(rf // 16) * h,
    aux_mode=1, aux_row_div=16, relu=True, a2=upper_c, lda2=h, a2_bs=rf * h, k1=4 * q), flops=2.0 * nb * rf * h * kc)
w2 = (torch.randn(h, h, device=dev) * 0.03).to(bf)
h2 = torch.empty(m, h, dtype=bf, device=dev)
timed('NT expand    m x 1024 x 1024', lambda: ops.gemm_nt(h1, w2, h2, m, h, h, h, h, h, bias=bias, relu=True),
      flops=2.0 * m * h * h)
timed('NT dgrad+gate m x 1024 x 1024', lambda: ops.gemm_nt(h1, w2, h2, m, h, h, h, h, h, aux=h1, ldaux=h, aux_mode=2),
      flops=2.0 * m * h * h)
dw = torch.zeros(h, 2 * h, device=dev)
timed('TN wgrad 1024 x 2048 x m', lambda: ops.gemm_tn(h1, cat, dw, h, 2 * h, m, h, 2 * h, 2 * h), flops=2.0 * m * h * 2 * h)
w3 = (torch.randn(q, h, device=dev) * 0.03).to(bf)
b3 = torch.zeros(q, device=dev)
tgt = torch.randint(0, 256, (m,), dtype=torch.uint8, device=dev)
lse = torch.empty(m, device=dev); lpt = torch.empty(m, device=dev)
timed('NLL fwd (logits stay in TMEM)', lambda: ops.gemm_nll(0, h2, w3, b3, tgt, m, h, h, h, lse=lse, logp_target=lpt),
      flops=2.0 * m * q * h)
rg = torch.full((m,), -1.0 / m, device=dev)
dl = torch.empty(m, q, dtype=bf, device=dev)
timed('NLL bwd (dlogits bf16)', lambda: ops.gemm_nll(2, h2, w3, b3, tgt, m, h, h, h, row_grad=rg, dlogits=dl),
      flops=2.0 * m * q * h)
x = (torch.rand(64 * 16015, device=dev) * 2 - 1) * 0.99
timed('quantize_ulaw 64 x 16015', lambda: ops.quantize_ulaw(x, want_i64=True, want_u8=True), nbytes=x.numel() * (4 + 8 + 1))
timed('colsum m x 1024 bf16', lambda: ops.colsum(h1, m, h, h), nbytes=m * h * 2)
idx = torch.randint(0, 256, (64, 16003), dtype=torch.uint8, device=dev)
timed('onehot_rows 64 x 16003 x 256', lambda: ops.onehot_rows(idx), nbytes=idx.numel() * 513)
# generation kernels (BASELINE config 5: 256 utterances)
bg = 256
hg1 = torch.randn(bg, h, device=dev).to(bf)
hg2 = torch.empty(bg, h, dtype=bf, device=dev)
timed('small-M NT 256 x 1024 x 1024 (cluster split-K 4)', lambda: ops.gemm_nt(hg1, w2, hg2, bg, h, h, h, h, h, bias=bias, relu=True),
      flops=2.0 * bg * h * h)
table = torch.randn(4 * q, h, device=dev).to(bf)
win = torch.randint(0, 256, (bg, 16), dtype=torch.uint8, device=dev)
pre = torch.randn(bg, 4, h, device=dev).to(bf)
timed('embed_sum 256 x (4 rows of 1024)', lambda: ops.embed_sum(table, win[:, 12:], 16, bg, 4, q, h, pre[:, 1], 4 * h, True, hg2, h),
      nbytes=bg * (5 * h * 2 + h * 2))
logits = torch.randn(bg, q, device=dev)
u = torch.rand(bg, device=dev)
picks = torch.empty(bg, dtype=torch.uint8, device=dev)
timed("log-softmax + draw 256 x 256", lambda: ops.sample_categorical(logits, bg, q, u, win, 16, picks, 1, normalise=True),
      nbytes=bg * q * 4)
