"""Where the time of a small-M GEMM launch goes (needs a build with SRNN_NVCC_EXTRA=-DSRNN_SMALL_TS): global-timer
stamps per CTA: 0 kernel entry, 1 set-up done, 2 first operand stage landed, 3 accumulator complete, 4 after the
cluster barrier, 5 slice reduced + stored, 6 exit."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import _lib, ops   # noqa: E402

bf = torch.bfloat16
m, n, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 1024, 1024)))
a = torch.randn(m, k, device='cuda').to(bf)
w = torch.randn(n, k, device='cuda').to(bf)
c = torch.empty(m, n, dtype=bf, device='cuda')
x = torch.empty(1 << 20, device='cuda')
for it in range(3):
    x.normal_()                                    # something else runs in between, like in the real step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm_nt(a, w, c, m, n, k, k, k, n)
    e1.record()
    torch.cuda.synchronize()
lib = _lib.load() if hasattr(_lib, 'load') else _lib.lib
buf = (ctypes.c_uint64 * (512 * 8))()
lib.srnn_debug_small_ts.argtypes = [ctypes.c_void_p]
assert lib.srnn_debug_small_ts(buf) == 0
ts = torch.tensor(list(buf), dtype=torch.int64).view(512, 8)
ncta = int((ts[:, 0] > 0).sum())
t0 = int(ts[:ncta, 0].min())
print(f'{m}x{n}x{k}: {ncta} CTAs, events {1e3 * e0.elapsed_time(e1):.1f} us')
for i, name in enumerate(['entry', 'setup done', 'first stage landed', 'accumulator done', 'after cluster barrier',
                          'reduced + stored', 'exit']):
    col = ts[:ncta, i] - t0
    print(f'{name:>24s}: min {int(col.min()):6d}  median {int(col.median()):6d}  max {int(col.max()):6d} ns')
