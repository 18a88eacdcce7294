"""Phase timeline of one rank's GEMM share with the T10_TRACE_PHASES debug build (tools/bin/libtor10_phases.so):
CTA entry -> first piece start -> main loops -> fix-up done.  python tools/trace_phases.py [world]
Build the debug library first (every .cu with -DT10_TRACE_PHASES, linked into tools/bin/libtor10_phases.so):
    for f in blockmap relayout permute_tma ggemm ggemm_tma ggemm_tf32 svd_small; do nvcc -DT10_TRACE_PHASES \
        -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --cudart shared -Iinclude \
        -Itor10_b200/csrc -c tor10_b200/csrc/$f.cu -o /tmp/$f.o; done
    nvcc -shared --cudart shared -o tools/bin/libtor10_phases.so /tmp/{blockmap,relayout,permute_tma,ggemm,ggemm_tma,ggemm_tf32,svd_small}.o"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TOR10_LIB"] = os.path.join(ROOT, "tools", "bin", "libtor10_phases.so")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from tor10_b200 import _lib
from tor10_b200 import _plan, parallel
import bench
from helpers import psym
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
A, B = bench.c3a_specs(4096)
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
if world > 1:
    parallel._STATE["transport"] = "none"
    plan = _plan.ContractPlan(a, b, shard=(0, world))
else:
    plan = _plan.get_contract_plan(a, b)
g = plan.gemm
Aar, Bar = a._arena_tensor(), b._arena_tensor()
Ap = torch.empty(plan.rel_a.packed.size, device=dev, dtype=torch.float64)
Cc = plan.Lo.new_arena(torch.float64, zero=True)
plan.rel_a.run(Aar, Ap)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
for _ in range(3): g.run(Ap, Bar, Cc)
torch.cuda.synchronize()
res = {}
for cold in (True, False):
    n = g.sk_npieces
    trace = torch.zeros((n, 4), dtype=torch.int64, device=dev)
    lib = _lib.load()
    if cold: flush.fill_(1.0)
    torch.cuda.synchronize()
    lib.t10_ggemm_set_trace(trace.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.run(Ap, Bar, Cc); e.record()
    torch.cuda.synchronize()
    lib.t10_ggemm_set_trace(None)
    tr = trace.cpu().numpy()
    done, entry, t0, t1 = tr[:, 0], tr[:, 1], tr[:, 2], tr[:, 3]
    ok = entry > 0
    base = entry[ok].min()
    slots = g.d_sk_slots.cpu().numpy()
    first = np.array([slots[c] for c in range(len(slots) - 1) if slots[c + 1] > slots[c]])
    last = np.array([slots[c + 1] - 1 for c in range(len(slots) - 1) if slots[c + 1] > slots[c]])
    res["cold_L2" if cold else "warm_L2"] = {
        "event_us": s.elapsed_time(e) * 1e3,
        "cta_entry_spread_us": float((entry[ok].max() - base) * 1e-3),
        "entry_to_first_piece_us_p50_max": [float(np.median(t0[first] - entry[first]) * 1e-3), float((t0[first] - entry[first]).max() * 1e-3)],
        "last_loop_end_us_p50_max": [float(np.median(t1[last] - base) * 1e-3), float((t1[last].max() - base) * 1e-3)],
        "last_piece_done_us_p50_max": [float(np.median(done[last] - base) * 1e-3), float((done[last].max() - base) * 1e-3)],
        "loop_end_to_done_us_p50_max": [float(np.median(done[last] - t1[last]) * 1e-3), float((done[last] - t1[last]).max() * 1e-3)]}
print(json.dumps({"world": world, **res}))
