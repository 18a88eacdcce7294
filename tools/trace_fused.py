"""Where the fused GEMM + all-gather kernel spends its time on a sharded rank (bench workload, torchrun):
per-piece timeline (t10_ggemm_set_trace) of the SAME stream-K launch with and without the peer pushes.
Prints, per variant: kernel duration by CUDA events, main-loop span, the gaps between consecutive pieces of a
CTA (epilogue: fix-up, staging, bulk-copy issue, waits), and the time between the last main loop and kernel end.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_fused.py
"""
import os, sys, json, ctypes as C
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import tor10_b200 as T
from tor10_b200 import _plan, _lib, parallel
import bench
from helpers import psym

A, B = bench.c3a_specs(4096)
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
parallel.enable(transport="fused")
_plan.clear_caches()
pl = _plan.get_contract_plan(a, b)
for _ in range(3):
    T.Contract(a, b)
torch.cuda.synchronize(); dist.barrier()
g, st, lib = pl.gemm, pl._stage, _lib.load()
dt = torch.float64
Aa, Bb = a._arena_tensor(), b._arena_tensor()
if pl.rel_a is not None:
    Aa = pl.rel_a.run(Aa, torch.empty(pl.rel_a.packed.size, device=dev, dtype=dt))
if pl.rel_b is not None:
    Bb = pl.rel_b.run(Bb, torch.empty(pl.rel_b.packed.size, device=dev, dtype=dt))
Cc = pl.Lo.new_arena(dt, zero=False)
npc = g.sk_npieces if g.streamk else g.ntiles
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)


def launch(variant):
    if variant == "local":
        g.run(Aa, Bb, Cc)
        return
    k = st.turn
    if variant == "fused":
        ptrs = (C.c_void_p * st.world)(*st.peer_ptrs[k]); fl = (C.c_void_p * st.world)(*st.peer_flag_ptrs[k])
    else:    # "self": the same epilogue, one destination: this rank's own staging arena
        ptrs = (C.c_void_p * 1)(st.peer_ptrs[k][rank]); fl = (C.c_void_p * 1)(st.peer_flag_ptrs[k][rank])
    st.steps[k] += 1
    # TOR10_TRACE_SELF_BULK=1: the local arena through the copy engine too (how r01_trace_fused_n2.json was taken)
    bulk_self = os.environ.get("TOR10_TRACE_SELF_BULK") == "1"
    g.run_allgather(Aa, Bb, ptrs, fl, st.steps[k], st.done_ctr,
                    self_index=-1 if bulk_self else (rank if variant == "fused" else 0))


out = {}
for variant in ("local", "self", "fused"):
    for _ in range(3):
        launch(variant)
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(10):
        flush.fill_(1.0)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(variant); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    tr = torch.zeros((npc, 4), dtype=torch.int64, device=dev)
    lib.t10_ggemm_set_trace(tr.data_ptr())
    flush.fill_(1.0); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(variant); e1.record(); torch.cuda.synchronize()
    lib.t10_ggemm_set_trace(None)
    t = tr.cpu().numpy()
    cta, t0, t1 = t[:, 1], t[:, 2], t[:, 3]
    ok = t1 > 0
    start = t0[ok].min()
    gaps, ends = [], []
    for c in np.unique(cta[ok]):
        m = ok & (cta == c)
        o = np.argsort(t0[m]); s_, e_ = t0[m][o], t1[m][o]
        gaps.append(float((s_[1:] - e_[:-1]).sum()) * 1e-3)
        ends.append(float(e_[-1] - start) * 1e-3)
    out[variant] = {"kernel_us_median": round(float(np.median(ts)), 1), "kernel_us_traced_run": round(e0.elapsed_time(e1) * 1e3, 1),
                    "pieces": int(ok.sum()), "main_loop_span_us": round(float(t1[ok].max() - start) * 1e-3, 1),
                    "main_loop_busy_per_cta_us": round(float((t1[ok] - t0[ok]).sum()) * 1e-3 / len(gaps), 1),
                    "gap_per_cta_us_mean": round(float(np.mean(gaps)), 1), "gap_per_cta_us_max": round(float(np.max(gaps)), 1),
                    "cta_last_main_loop_end_us_median": round(float(np.median(ends)), 1)}
if rank == 0:
    print(json.dumps({"world": world, "tile_cfg": g.tile_cfg, "stream_k": bool(g.streamk), **out}))
dist.destroy_process_group()
