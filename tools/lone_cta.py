"""Main-loop efficiency of the fp64 GEMM kernels with 1, 2, 3 ... tiles per SM (long K, no tail effects):
M = 148 * BM rows, N = j * BN columns, K = 4096 -> exactly j tiles per SM.  Reports TFLOP/s per tile_cfg.
    python tools/lone_cta.py
"""
import os, sys, json, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tor10_b200 as T
from tor10_b200 import _plan, _lib

dev = torch.device("cuda", 0)
lib = _lib.load()


def timed(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3


K = int(os.environ.get("K", "4096"))
cfgs = [int(x) for x in os.environ.get("CFGS", "3,0,2,1,40,42").split(",")]
for cfg in cfgs:
    bm, bn = C.c_int(0), C.c_int(0)
    lib.t10_ggemm_tile_shape(cfg, C.byref(bm), C.byref(bn))
    row = {"cfg": cfg, "tile": "%dx%d" % (bm.value, bn.value)}
    for j in (1, 2, 3, 4, 6):
        M, N = 148 * bm.value, j * bn.value
        A = torch.randn((M, K), device=dev, dtype=torch.float64)
        B = torch.randn((K, N), device=dev, dtype=torch.float64)
        Cc = torch.empty((M, N), device=dev, dtype=torch.float64)
        prob = np.zeros(1, dtype=_lib.GEMM_PROBLEM_DTYPE)
        prob[0] = (0, 0, 0, M, N, K, K, N, N)
        grp = _plan.GemmGroup(prob, dev, torch.float64, tile_cfg=cfg, base_aligned=True)
        us = timed(lambda: grp.run(A, B, Cc))
        row["%d/SM" % j] = round(2.0 * M * N * K / us * 1e-6, 2)
        if cfg == cfgs[0]:
            us2 = timed(lambda: torch.matmul(A, B, out=Cc))
            row["cublas %d/SM" % j] = round(2.0 * M * N * K / us2 * 1e-6, 2)
    print(json.dumps(row), flush=True)
