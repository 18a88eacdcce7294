"""Per-piece timeline of ONE RANK's share of the bench contraction, run on a single GPU (the plan of rank r of N is
built without a process group): where a small share's time goes.   python tools/trace_share.py [world] [rank]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from tor10_b200 import _plan, _lib, parallel
import bench
from helpers import psym
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
A, B = bench.c3a_specs(4096)
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
parallel._STATE["transport"] = "none"
plan = _plan.ContractPlan(a, b, shard=(rank, world))
g = plan.gemm
Aar, Bar = a._arena_tensor(), b._arena_tensor()
Ap = torch.empty(plan.rel_a.packed.size, device=dev, dtype=torch.float64)
Cc = plan.Lo.new_arena(torch.float64, zero=True)
plan.rel_a.run(Aar, Ap)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
for _ in range(3): g.run(Ap, Bar, Cc)
torch.cuda.synchronize()
ev = []
for _ in range(20):
    flush.fill_(1.0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.run(Ap, Bar, Cc); e.record(); ev.append((s, e))
torch.cuda.synchronize()
us = float(np.mean([s.elapsed_time(e) for s, e in ev])) * 1e3
n = g.sk_npieces
trace = torch.zeros((n, 4), dtype=torch.int64, device=dev)
lib = _lib.load()
flush.fill_(1.0)
lib.t10_ggemm_set_trace(trace.data_ptr())
g.run(Ap, Bar, Cc)
torch.cuda.synchronize()
lib.t10_ggemm_set_trace(None)
tr = trace.cpu().numpy()
pcs = g.d_sk_pieces.cpu().numpy(); slots = g.d_sk_slots.cpu().numpy()
t0, t1 = tr[:, 2], tr[:, 3]
start = t0.min()
ks = pcs[:, 1] - pcs[:, 0]
per_cta_first = np.array([t0[slots[c]] for c in range(len(slots) - 1) if slots[c + 1] > slots[c]])
per_cta_last = np.array([t1[slots[c + 1] - 1] for c in range(len(slots) - 1) if slots[c + 1] > slots[c]])
out = {"world": world, "rank": rank, "event_us": us, "pieces": int(n), "parked": int(g.sk_parked), "ctas": int(len(per_cta_first)),
       "ksteps_total": int(ks.sum()), "main_loop_span_us": float((t1.max() - start) * 1e-3),
       "first_piece_start_us_p50_max": [float(np.median(per_cta_first - start) * 1e-3), float((per_cta_first.max() - start) * 1e-3)],
       "last_piece_loop_end_us_p10_p50_max": [float(np.percentile(per_cta_last - start, 10) * 1e-3), float(np.median(per_cta_last - start) * 1e-3), float((per_cta_last.max() - start) * 1e-3)],
       "us_per_kstep_in_loop": float(((t1 - t0) * 1e-3).sum() / ks.sum()),
       "share_of_flops": g.flops / plan.flops_total}
print(json.dumps(out))
