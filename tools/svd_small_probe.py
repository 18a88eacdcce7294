"""Duration and accuracy of the one-CTA Jacobi SVD against torch / cuSOLVER drivers."""
import json, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tor10_b200 import linalg
dev = torch.device("cuda", 0)
out = []
for shape in ((64, 64), (32, 32), (128, 64), (16, 16)):
    x = torch.rand(shape, device=dev, dtype=torch.float64)
    for _ in range(3): linalg._svd_small(x)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for s, e in ev:
        s.record(); u, sv, vh = linalg._svd_small(x); e.record()
    torch.cuda.synchronize()
    us = float(np.median([s.elapsed_time(e) for s, e in ev])) * 1e3
    ref = torch.linalg.svdvals(x)
    t0 = time.perf_counter()
    for _ in range(10): torch.linalg.svd(x, full_matrices=False)
    torch.cuda.synchronize(); cus = (time.perf_counter() - t0) / 10 * 1e6
    out.append({"shape": shape, "jacobi_us": us, "cusolver_default_us": cus,
                "sv_rel_err": float(((sv - ref).abs().max() / ref.max()).item()),
                "recon_err": float((((u * sv) @ vh - x).abs().max()).item())})
# sweeps per matrix (info word) for a random matrix and for an iTEBD-like one (fast-decaying singular values)
import ctypes as C
from tor10_b200 import _lib
def sweeps(x):
    m, n = x.shape
    u = torch.empty((m, n), device=dev, dtype=x.dtype); sv = torch.empty(n, device=dev, dtype=x.dtype); vh = torch.empty((n, n), device=dev, dtype=x.dtype)
    info = torch.zeros(1, device=dev, dtype=torch.int32)
    prob = (C.c_int64 * 9)(m, n, 0, x.stride(0), x.stride(1), 0, 0, 0, -1)
    _lib.check(_lib.load().t10_svd_small(0, 1, prob, x.data_ptr(), u.data_ptr(), sv.data_ptr(), vh.data_ptr(), None, info.data_ptr(), _lib.stream_ptr()))
    v = int(info.item()); return {"sweeps": v >> 1, "not_converged": v & 1}
x = torch.rand((64, 64), device=dev, dtype=torch.float64)
q1, _ = torch.linalg.qr(torch.rand((64, 64), device=dev, dtype=torch.float64)); q2, _ = torch.linalg.qr(torch.rand((64, 64), device=dev, dtype=torch.float64))
decay = (q1 * torch.logspace(0, -40, 64, device=dev, dtype=torch.float64)) @ q2
out.append({"random": sweeps(x), "decaying_1e-40": sweeps(decay), "rank8": sweeps((x[:, :8] @ x[:8, :]).contiguous())})
print(json.dumps(out))
