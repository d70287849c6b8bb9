"""Does a bf16 NaN in one element of A[m, k] reach every column of row m of D = A . B^T on the tcgen05 tensor cores
(and only that row)?  The recurrent kernel's fence-free exchange validates its operand this way."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import ops   # noqa: E402

m, n, k = 256, 1024, 1024
a = torch.randn(m, k, device='cuda').to(torch.bfloat16)
b = torch.zeros(n, k, device='cuda', dtype=torch.bfloat16)          # zeros: NaN * 0 must still be NaN
b[::2] = torch.randn(n // 2, k, device='cuda').to(torch.bfloat16)
a.view(torch.int16)[7, 513] = -1                                     # 0xFFFF: the sentinel pattern
c = torch.empty(m, n, device='cuda', dtype=torch.float32)
ops.gemm_nt(a, b, c, m, n, k, k, k, n)
torch.cuda.synchronize()
bad_rows = torch.isnan(c).any(dim=1).nonzero().flatten().tolist()
print('rows with NaN:', bad_rows, ' all columns of row 7 NaN:', bool(torch.isnan(c[7]).all()))
assert bad_rows == [7] and bool(torch.isnan(c[7]).all())
print('nan probe ok')
