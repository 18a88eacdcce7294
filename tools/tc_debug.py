import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TOR10_F32_CFG"] = "50"
import tor10_b200 as T
from tor10_b200.UniTensor import _dense_gemm
dev = torch.device("cuda", 0)
def run(A, B, name):
    C = _dense_gemm(A.contiguous(), B.contiguous())
    torch.cuda.synchronize()
    want = A.double() @ B.double()
    err = (C.double() - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
    print(name, "err", err, "C[0,:6]", C[0, :6].tolist(), "want", want[0, :6].tolist())
    return C, want
M = N = 128
for K in (8, 32, 64):
    A = torch.ones((M, K), device=dev, dtype=torch.float32); B = torch.ones((K, N), device=dev, dtype=torch.float32)
    run(A, B, "ones K=%d" % K)
K = 32
A = torch.zeros((M, K), device=dev, dtype=torch.float32); A[:, 0] = 1.0
B = torch.arange(K * N, device=dev, dtype=torch.float32).reshape(K, N) / 64.0
C, W = run(A, B, "A[:,0]=1 (C rows = B row 0)")
A = torch.zeros((M, K), device=dev, dtype=torch.float32); A[:, 5] = 1.0
C, W = run(A, B, "A[:,5]=1 (C rows = B row 5)")
print(" C[3,:8]", C[3, :8].tolist(), "\n W[3,:8]", W[3, :8].tolist())
A = torch.arange(M * K, device=dev, dtype=torch.float32).reshape(M, K) / 64.0
B = torch.zeros((K, N), device=dev, dtype=torch.float32); B[2, :] = 1.0
C, W = run(A, B, "B[2,:]=1 (C[:, j] = A[:, 2])")
print(" C[:8,0]", C[:8, 0].tolist(), "\n W[:8,0]", W[:8, 0].tolist())
g = torch.Generator().manual_seed(0)
A = torch.rand((M, 64), generator=g).to(dev); B = torch.rand((64, N), generator=g).to(dev)
C, W = run(A, B, "random 128x128x64")
bad = ((C.double() - W).abs() > 1e-3).nonzero()
print(" mismatches", bad.shape[0], bad[:10].tolist())
flag = ctypes.c_int(0); T._lib.load().t10_ggemm_tf32_status(ctypes.byref(flag)); print("barrier flag", flag.value)
