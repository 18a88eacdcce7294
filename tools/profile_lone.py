"""One fp64 GEMM with exactly J tiles per SM (long K) through tile_cfg CFG, for ncu source-level stalls.
    CFG=3 J=1 python tools/profile_lone.py
"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tor10_b200 as T
from tor10_b200 import _plan, _lib
dev = torch.device("cuda", 0)
cfg, J, K = int(os.environ.get("CFG", "3")), int(os.environ.get("J", "1")), int(os.environ.get("K", "4096"))
bm, bn = C.c_int(0), C.c_int(0)
_lib.load().t10_ggemm_tile_shape(cfg, C.byref(bm), C.byref(bn))
M, N = 148 * bm.value, J * bn.value
A = torch.randn((M, K), device=dev, dtype=torch.float64)
B = torch.randn((K, N), device=dev, dtype=torch.float64)
Cc = torch.empty((M, N), device=dev, dtype=torch.float64)
prob = np.zeros(1, dtype=_lib.GEMM_PROBLEM_DTYPE)
prob[0] = (0, 0, 0, M, N, K, K, N, N)
grp = _plan.GemmGroup(prob, dev, torch.float64, tile_cfg=cfg, base_aligned=True)
for _ in range(3):
    grp.run(A, B, Cc)
torch.cuda.synchronize()
print("ok")
