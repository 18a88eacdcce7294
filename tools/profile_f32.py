"""One dense fp32 4096^3 GEMM through the tcgen05 3xTF32 kernel (for `ncu -k regex:tf32x3`)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tor10_b200 as T
from tor10_b200 import _plan, _lib
dev = torch.device("cuda", 0)
M = N = K = 4096
A = torch.randn((M, K), device=dev, dtype=torch.float32)
B = torch.randn((K, N), device=dev, dtype=torch.float32)
C = torch.empty((M, N), device=dev, dtype=torch.float32)
prob = np.zeros(1, dtype=_lib.GEMM_PROBLEM_DTYPE)
prob[0] = (0, 0, 0, M, N, K, K, N, N)
grp = _plan.GemmGroup(prob, dev, torch.float32, tile_cfg=50, base_aligned=True)
for _ in range(3):
    grp.run(A, B, C)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
