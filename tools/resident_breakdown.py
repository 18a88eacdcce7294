"""Where a rank's time goes in a sector-resident Contract of the bench workload (torchrun, N ranks): relayout of its
share, GEMM of its share with and without the progress signal, the whole resident Contract, and the host time per call.
CUDA events, L2 flushed between calls, max over ranks."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
import tor10_b200 as T
from tor10_b200 import _plan, parallel, _lib
import bench
from helpers import psym
parallel.enable(transport=os.environ.get("TOR10_GATHER", "resident"))
A, B = bench.c3a_specs(4096)
a, b = psym(T, A, seed=1003, device=dev), psym(T, B, seed=1004, device=dev)
plan = _plan.get_contract_plan(a, b)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
def timed(fn, n=30, flush_l2=True):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    evs = []
    for _ in range(n):
        if flush_l2: flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    t = torch.tensor([float(np.mean([s.elapsed_time(e) for s, e in evs]))], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3
Aar, Bar = a._arena_tensor(), b._arena_tensor()
Ap = torch.empty(plan.rel_a.packed.size, device=dev, dtype=torch.float64)
Cc = plan.Lo.new_arena(torch.float64, zero=True)
plan.rel_a.run(Aar, Ap)
rw = parallel.resident_world(dev)
out = {"world": world, "transport": plan._transport, "tile_cfg": plan.gemm.tile_cfg, "tma": bool(plan.gemm.tma),
       "pieces": int(getattr(plan.gemm, "sk_npieces", 0)), "parked": int(getattr(plan.gemm, "sk_parked", 0)),
       "share_of_flops": plan.gemm.flops / plan.flops_total}
out["relayout_us"] = timed(lambda: plan.rel_a.run(Aar, Ap))
out["gemm_us"] = timed(lambda: plan.gemm.run(Ap, Bar, Cc))
out["gemm_signal_us"] = timed(lambda: plan.gemm.run(Ap, Bar, Cc, signal=(rw.sig_ptrs, rw.next_seq(), rw.done_ctr)))
out["relayout_gemm_pdl_us"] = timed(lambda: (plan.rel_a.run(Aar, Ap), plan.gemm.run(Ap, Bar, Cc, pdl=True)))
keep = [None]
def step(): keep[0] = T.Contract(a, b)
out["contract_us"] = timed(step)
out["contract_noflush_us"] = timed(step, flush_l2=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter(); torch.cuda.synchronize()
out["host_us_per_call"] = (t1 - t0) / 200 * 1e6
# kernel timeline of a few steps (Kineto): names, durations and the gaps between consecutive kernels of a step
try:
    from torch.profiler import profile, ProfilerActivity
    for _ in range(5): step()
    torch.cuda.synchronize(); dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(6):
            flush.fill_(1.0); step()
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    tl, prev_end = [], None
    for e in evs:
        st, en = e.time_range.start, e.time_range.end
        tl.append({"name": e.name[:60], "us": round(en - st, 1), "gap_before_us": None if prev_end is None else round(st - prev_end, 1)})
        prev_end = en
    out["timeline"] = tl[-24:]
except Exception as exc:
    out["timeline_error"] = repr(exc)[:200]
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
