"""Call-to-call variance of batched generation (diagnostic): python scripts/gen_variance.py [--frames 4000]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from samplernn_pase_b200 import SampleRNNModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=4000)
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--reps', type=int, default=4)
a = ap.parse_args()
torch.manual_seed(0)
model = SampleRNNModel(**bench.model_kwargs(a.frames)).cuda()
utt = torch.randn(a.batch, a.frames, 43).cuda()
info = [{'speaker': {'index': i % 126}} for i in range(a.batch)]
model.test(utt[:, :3], info)
torch.cuda.synchronize()
for frames in (a.frames // 4, a.frames // 2, a.frames):
    for r in range(a.reps):
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.test(utt[:, :frames], info)
        e1.record()
        t_host = time.perf_counter() - t0                 # host returns when everything is enqueued
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f'frames {frames:5d} rep {r}: host enqueue {t_host:.3f} s, total {t_all:.3f} s, device {e0.elapsed_time(e1) / 1e3:.3f} s '
              f'-> {1e6 * t_all / (frames * 16):.1f} us per sample step', flush=True)
