"""Generation: host cost vs device time of one CUDA-graph replay (GRAPH_FRAMES top-tier frames).
python scripts/gen_hostdev.py [--batch 256] [--frames 400] [--graph-frames 8] [--no-pdl]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from samplernn_pase_b200 import SampleRNNModel, generate  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--frames', type=int, default=400)
ap.add_argument('--graph-frames', type=int, default=8)
ap.add_argument('--no-pdl', action='store_true')
a = ap.parse_args()
generate.GRAPH_FRAMES = a.graph_frames
generate._PDL = not a.no_pdl
torch.manual_seed(0)
model = SampleRNNModel(**bench.model_kwargs(a.frames)).cuda()
utt = torch.randn(a.batch, a.frames, 43).cuda()
info = [{'speaker': {'index': i % 126}} for i in range(a.batch)]
model.test(utt[:, :3], info)
torch.cuda.synchronize()

host, events = [], []
orig = torch.cuda.CUDAGraph.replay


def replay(self):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    orig(self)
    host.append(time.perf_counter() - t0)
    e1.record()
    events.append((e0, e1))


torch.cuda.CUDAGraph.replay = replay
t0 = time.perf_counter()
model.test(utt, info)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
fs = int(model.frame_size)
steps = a.graph_frames * fs
dev = sorted(e0.elapsed_time(e1) for e0, e1 in events)
hst = sorted(1e3 * h for h in host)
med = lambda v: v[len(v) // 2]
print(f'batch {a.batch}, {a.frames} frames, {len(events)} replays of {steps} sample steps, PDL {"off" if a.no_pdl else "on"}: '
      f'device {med(dev):.3f} ms per replay (min {dev[0]:.3f}, max {dev[-1]:.3f}) = {1e3 * med(dev) / steps:.1f} us per sample step; '
      f'host cudaGraphLaunch {med(hst):.3f} ms per replay (min {hst[0]:.3f}, max {hst[-1]:.3f}) = {1e3 * med(hst) / steps:.1f} us per sample step; '
      f'whole call {wall:.3f} s = {1e6 * wall / (a.frames * fs):.1f} us per sample step')
