"""fp32 grouped GEMM: tcgen05 3xTF32 kernel (cfg 50) against the SIMT kernel (cfg 0) and cuBLAS fp32 / TF32.
Accuracy (relative to an fp64 product, normalised by max|C|) and time on dense and block-sparse workloads.
    python tools/bench_f32.py
"""
import os, sys, json, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tor10_b200 as T
from tor10_b200 import _plan, _lib

dev = torch.device("cuda", 0)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3   # us


def group(M, N, K, cfg):
    prob = np.zeros(1, dtype=_lib.GEMM_PROBLEM_DTYPE)
    prob[0] = (0, 0, 0, M, N, K, K, N, N)
    return _plan.GemmGroup(prob, dev, torch.float32, tile_cfg=cfg, base_aligned=True)


def dense(M, N, K, positive):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.rand((M, K), generator=g) if positive else torch.randn((M, K), generator=g)
    B = torch.rand((K, N), generator=g) if positive else torch.randn((K, N), generator=g)
    A, B = A.to(dev), B.to(dev)
    want = A.double() @ B.double()
    scale = want.abs().max().item()
    out = {"case": "dense %dx%dx%d %s" % (M, N, K, "positive" if positive else "gaussian")}
    flops = 2.0 * M * N * K
    for name, cfg in (("tcgen05_3xtf32", 50), ("simt", 0)):
        grp = group(M, N, K, cfg)
        C_ = torch.empty((M, N), device=dev, dtype=torch.float32)
        grp.run(A, B, C_)
        torch.cuda.synchronize()
        err = (C_.double() - want).abs().max().item() / scale
        us = timed(lambda: grp.run(A, B, C_))
        out[name] = {"err": err, "us": us, "tflops": flops / us * 1e-6}
    for name, allow in (("cublas_fp32", False), ("cublas_tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = allow
        C_ = A @ B
        err = (C_.double() - want).abs().max().item() / scale
        us = timed(lambda: torch.matmul(A, B))
        out[name] = {"err": err, "us": us, "tflops": flops / us * 1e-6}
    torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out), flush=True)


def symmetric(chi):
    """Config-3a shaped block-sparse contraction in fp32 (same bond structure as bench.py's workload)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from helpers import psym
    A, B = bench.c3a_specs(chi)
    a = psym(T, A, seed=1003, device=dev, dtype=torch.float32)
    b = psym(T, B, seed=1004, device=dev, dtype=torch.float32)
    res = {"gflop": _plan.get_contract_plan(a, b).flops_total * 1e-9}
    for name, cfg in (("tcgen05_3xtf32", "50"), ("simt", "0")):
        os.environ["TOR10_F32_CFG"] = cfg
        _plan.clear_caches()
        c = T.Contract(a, b)
        torch.cuda.synchronize()
        us = timed(lambda: T.Contract(a, b), iters=10)
        res[name] = {"us": us}
    print(json.dumps({"case": "C3a chi=%d fp32 (whole Contract)" % chi, **res}), flush=True)


if __name__ == "__main__":
    for (M, N, K) in ((4096, 4096, 4096), (1024, 1024, 8192), (512, 512, 512), (300, 200, 100)):
        for positive in (True, False):
            dense(M, N, K, positive)
    symmetric(4096)
    flag = ctypes.c_int(0)
    _lib.load().t10_ggemm_tf32_status(ctypes.byref(flag))
    print("tcgen05 barrier flag", flag.value)
