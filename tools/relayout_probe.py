"""Relayout replay kernels on the bench operand: event-timed after an L2 flush and back-to-back (no flush)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from tor10_b200 import _plan
import bench
from helpers import psym
dev = torch.device("cuda", 0)
A, B = bench.c3a_specs(int(os.environ.get("CHI", "4096")))
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
for mode, env in (("runs", {"TOR10_RELAYOUT_RUNS": "1", "TOR10_RELAYOUT_INDEX": "1"}),
                  ("index", {"TOR10_RELAYOUT_RUNS": "0", "TOR10_RELAYOUT_INDEX": "1"}),
                  ("blocks", {"TOR10_RELAYOUT_RUNS": "0", "TOR10_RELAYOUT_INDEX": "0"})):
    os.environ.update(env)
    _plan.clear_caches()
    pl = _plan.get_contract_plan(a, b)
    rel = pl.rel_a
    src = a._arena_tensor()
    dst = torch.empty(rel.packed.size, device=dev, dtype=torch.float64)
    for _ in range(3):
        rel.run(src, dst)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); rel.run(src, dst); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        rel.run(src, dst)
    e.record(); torch.cuda.synchronize()
    warm = s.elapsed_time(e) * 1e3 / 50
    nbytes = 16.0 * rel.moved_elems
    print(json.dumps({"mode": mode, "cold_us": float(np.median(ts)), "back_to_back_us": warm,
                      "cold_GBs": nbytes / np.median(ts) * 1e-3, "b2b_GBs": nbytes / warm * 1e-3,
                      "nruns": getattr(rel, "nruns", None) if rel.d_runs is not None else None}), flush=True)
