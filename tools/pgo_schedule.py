"""Profile-guided tile schedule experiment: re-run LPT on MEASURED per-tile durations (t10_ggemm_set_trace)
instead of the DMMA-count model, iterate, and time the GEMM stage after every round.
    python tools/pgo_schedule.py [rounds]
"""
import os, sys, heapq, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from tor10_b200 import _plan, _lib
import bench
from helpers import psym

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
A, B = bench.c3a_specs(4096)
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
plan = _plan.get_contract_plan(a, b)
g = plan.gemm
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
Aa, Bb = a._arena_tensor(), b._arena_tensor()
Ap = torch.empty(plan.rel_a.packed.size, device=dev, dtype=torch.float64) if plan.rel_a is not None else Aa
if plan.rel_a is not None:
    plan.rel_a.run(Aa, Ap)
Cc = plan.Lo.new_arena(torch.float64, zero=True)


def time_gemm(n=40):
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.run(Ap, Bb, Cc); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.mean(ts)), float(np.min(ts))


def trace_once():
    tr = torch.zeros((g.ntiles, 4), dtype=torch.int64, device=dev)
    lib.t10_ggemm_set_trace(tr.data_ptr())
    flush.fill_(1.0)
    g.run(Ap, Bb, Cc)
    torch.cuda.synchronize()
    lib.t10_ggemm_set_trace(None)
    t = tr.cpu().numpy()
    return (t[:, 3] - t[:, 2]) * 1e-3, t


def install(order_slots):
    """order_slots: list of lists of tile rows (each [p, m0, n0, 0])."""
    tiles = np.concatenate([np.array(s_, dtype=np.int32).reshape(-1, 4) for s_ in order_slots], axis=0)
    starts = np.zeros(len(order_slots) + 1, dtype=np.int32)
    starts[1:] = np.cumsum([len(s_) for s_ in order_slots])
    g.d_tiles = torch.from_numpy(np.ascontiguousarray(tiles)).to(dev)
    g.d_slots = torch.from_numpy(starts).to(dev)


print("cfg", g.tile_cfg, "tiles", g.ntiles, "slots", g.nslots)
for _ in range(3):
    g.run(Ap, Bb, Cc)
torch.cuda.synchronize()
print(json.dumps({"round": 0, "gemm_us_mean_min": time_gemm()}))
for r in range(1, rounds + 1):
    durs = np.mean([trace_once()[0] for _ in range(3)], axis=0)
    tiles = g.d_tiles.cpu().numpy()[: g.ntiles]
    order = np.argsort(-durs, kind="stable")
    heap = [(0.0, s_) for s_ in range(g.nslots)]
    heapq.heapify(heap)
    slots = [[] for _ in range(g.nslots)]
    for i in order:
        load, s_ = heapq.heappop(heap)
        slots[s_].append(tiles[i])
        heapq.heappush(heap, (load + float(durs[i]), s_))
    install(slots)
    for _ in range(3):
        g.run(Ap, Bb, Cc)
    torch.cuda.synchronize()
    loads = sorted(l for l, _ in heap)
    print(json.dumps({"round": r, "gemm_us_mean_min": time_gemm(), "model_makespan_over_mean": loads[-1] / np.mean(loads)}))
