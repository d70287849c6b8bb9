"""Co-residency of thread-block clusters on this GPU (cudaOccupancyMaxActiveClusters through srnn_probe_clusters)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401
from samplernn_pase_b200 import _lib  # noqa: E402

torch.cuda.init()
for cs in (2, 4, 8, 16):
    for smem in (64, 128, 160, 200, 224):
        n = C.c_int32(0)
        try:
            _lib.call('srnn_probe_clusters', cs, smem * 1024, 256, C.byref(n))
            print(f'cluster {cs:2d} x {smem:3d} KB smem, 256 threads: {n.value:3d} clusters = {n.value * cs:3d} CTAs co-resident')
        except RuntimeError as e:
            print(f'cluster {cs:2d} x {smem:3d} KB: {str(e)[-100:]}')
