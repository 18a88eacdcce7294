#!/usr/bin/env python
"""Wall-clock cost per call of small operations (host-side overhead of the Python layer + launches)."""
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T  # noqa: E402

dev = torch.device("cuda", 0)


def per_call(fn, n=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return {"host_us": (t1 - t0) / n * 1e6, "with_gpu_drain_us": (t2 - t0) / n * 1e6}


chi = 32
A = T.UniTensor(bonds=[T.Bond(chi), T.Bond(2), T.Bond(chi)], rowrank=1, labels=[-1, 0, -2], device=dev).Rand()
B = T.UniTensor(bonds=[T.Bond(chi), T.Bond(2), T.Bond(chi)], rowrank=1, labels=[-3, 1, -4], device=dev).Rand()
la = T.UniTensor(bonds=[T.Bond(chi), T.Bond(chi)], rowrank=1, labels=[-2, -3], is_diag=True, device=dev).Rand()
X = T.UniTensor(bonds=[T.Bond(chi), T.Bond(2), T.Bond(2), T.Bond(chi)], rowrank=1, labels=[-4, 0, 1, -5], device=dev).Rand()
H = T.UniTensor(bonds=[T.Bond(2)] * 4, rowrank=2, labels=[0, 1, 2, 3], device=dev).Rand()
Ala = T.Contract(A, la)
M = T.UniTensor(bonds=[T.Bond(2 * chi), T.Bond(2 * chi)], rowrank=1, device=dev).Rand()
res = {
    "Contract(A, diag)": per_call(lambda: T.Contract(A, la)),
    "Contract(Ala, B) [chi,2,chi]x[chi,2,chi]": per_call(lambda: T.Contract(Ala, B)),
    "Contract(X, H)": per_call(lambda: T.Contract(X, H)),
    "Contract(X, X) full": per_call(lambda: T.Contract(X, X)),
    "deepcopy(X)": per_call(lambda: copy.deepcopy(X)),
    "Svd_truncate 64x64": per_call(lambda: T.Svd_truncate(M, chi), n=200),
    "torch.linalg.svd 64x64": per_call(lambda: torch.linalg.svd(M.Storage, full_matrices=False), n=200),
    "Inverse(diag)": per_call(lambda: T.Inverse(la)),
    "Reshape": per_call(lambda: A.Reshape([chi * 2, chi], rowrank=1)),
    "empty kernel-ish: torch.empty": per_call(lambda: torch.empty(16, device=dev)),
}
print(json.dumps(res, indent=1))
