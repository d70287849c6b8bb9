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
a = ap.parse_args()
torch.manual_seed(0)
model = SampleRNNModel(**bench.model_kwargs(a.frames)).cuda()
utt = torch.randn(a.batch, a.frames, 43).cuda()
info = [{'speaker': {'index': i % 126}} for i in range(a.batch)]
model.test(utt[:, :2], info)                   # warm-up
torch.cuda.synchronize()
ops.launch_count = 0
t0 = time.perf_counter()
y = model.test(utt, info)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n = a.frames * int(model.frame_size)
print(f'generated {a.batch} x {n} samples in {dt:.3f} s: {1e6 * dt / n:.1f} us per sample step, '
      f'{a.batch * n / dt / 1e3:.1f} k samples/s, {ops.launch_count / n:.1f} kernel launches per sample step')
