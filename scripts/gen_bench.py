"""Generation throughput (BASELINE config 5 shape: C2 model, 256 utterances) on a short horizon.
python scripts/gen_bench.py [--batch 256] [--frames 8]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from samplernn_pase_b200 import SampleRNNModel, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--frames', type=int, default=8)
ap.add_argument('--profile', action='store_true')
ap.add_argument('--eager', action='store_true')
ap.add_argument('--gru-flags', type=int, default=0)
a = ap.parse_args()
torch.manual_seed(0)
ops.gru_tuning_flags = a.gru_flags
model = SampleRNNModel(**bench.model_kwargs(a.frames)).cuda()
utt = torch.randn(a.batch, a.frames, 43).cuda()
info = [{'speaker': {'index': i % 126}} for i in range(a.batch)]
model.test(utt[:, :3], info)                   # warm-up (lazy initialisation, allocator)
torch.cuda.synchronize()


def run(frames):
    ops.launch_count = 0
    t0 = time.perf_counter()
    model.test(utt[:, :frames], info, use_graphs=not a.eager)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


# the call prepares the weights and captures the step graphs once: report the whole call and the marginal step
short = max(a.frames // 4, 3)
run(short)                                     # untimed: the first graph capture of a process pays one-off costs
dt_short, dt = run(short), run(a.frames)
fs = int(model.frame_size)
n = a.frames * fs
step_us = 1e6 * (dt - dt_short) / ((a.frames - short) * fs)
print(f'[{short} frames {dt_short:.3f} s, {a.frames} frames {dt:.3f} s] generated {a.batch} x {n} samples in {dt:.3f} s (whole call: {a.batch * n / dt / 1e3:.1f} k samples/s); '
      f'marginal {step_us:.1f} us per sample step = {a.batch / step_us * 1e3:.1f} k samples/s; '
      f'setup + graph capture {1e3 * (dt - step_us * n * 1e-6):.1f} ms')
if '--profile' in sys.argv:
    # device time of every C-ABI call of the eager path (CUDA events; launch gaps excluded)
    import collections
    from samplernn_pase_b200 import _lib
    _lib.profile_log = []
    model.test(utt, info, use_graphs=False)
    torch.cuda.synchronize()
    log, _lib.profile_log = _lib.profile_log, None
    agg = collections.OrderedDict()
    for name, note, s, e in log:
        d = agg.setdefault((name, note if 'gemm' in name or 'gru' in name else ''), [0, 0.0])
        d[0] += 1
        d[1] += s.elapsed_time(e)
    tot = sum(v[1] for v in agg.values())
    print(f'eager path: {len(log) / n:.1f} C calls and {1e3 * tot / n:.1f} us of device time per sample step')
    for (name, note), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{1e3 * ms / n:8.2f} us/sample  x{cnt / n:5.2f}  {1e3 * ms / cnt:7.2f} us each  {name:22s} {note}')
