"""Stage-by-stage device time of a sector-sharded Contract (bench workload) under torchrun:
relayout A / relayout B / result arena / GEMM(+fused all-gather stores) / barrier / copy-out, per transport.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_breakdown.py
"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import tor10_b200 as T
from tor10_b200 import _plan, parallel
import bench
from helpers import psym

A, B = bench.c3a_specs(int(os.environ.get("CHI", "4096")))
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)
dt = torch.float64


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(transport, iters=30):
    parallel.enable(transport=transport)
    _plan.clear_caches()
    pl = _plan.get_contract_plan(a, b)
    for _ in range(3):
        T.Contract(a, b)
    torch.cuda.synchronize(); dist.barrier()
    names = ["relayout_a", "relayout_b", "arena", "gemm", "barrier", "copy_out", "total"]
    acc = {k: [] for k in names}
    for _ in range(iters):
        flush.fill_(1.0)
        dist.barrier(); torch.cuda.synchronize()
        e = [ev() for _ in range(7)]
        Aa, Bb = a._arena_tensor(), b._arena_tensor()
        e[0].record()
        if pl.rel_a is not None:
            Aa = pl.rel_a.run(Aa, torch.empty(pl.rel_a.packed.size, device=dev, dtype=dt))
        e[1].record()
        if pl.rel_b is not None:
            Bb = pl.rel_b.run(Bb, torch.empty(pl.rel_b.packed.size, device=dev, dtype=dt))
        e[2].record()
        fused = pl._transport == "fused" and pl.gemm.fusable()
        Cc = pl.Lo.new_arena(dt, zero=(not fused) and (not pl.all_matched or pl._transport == "none"))
        e[3].record()
        if pl._stage is None and pl._transport in ("fused", "p2p"):
            pl._stage = parallel.symm_staging(pl, Cc.numel(), dt, dev)
        st = pl._stage
        if fused:
            import ctypes as C
            k = st.turn
            ptrs = (C.c_void_p * st.world)(*st.peer_ptrs[k])
            fptrs = (C.c_void_p * st.world)(*st.peer_flag_ptrs[k])
            st.steps[k] += 1
            pl.gemm.run_allgather(Aa, Bb, ptrs, fptrs, st.steps[k], st.done_ctr, self_index=st.rank)
            e[4].record()
            e[5].record()
            n = Cc.numel() // 4 * 4
            T._lib.check(T._lib.load().t10_wait_copy(T._lib.dtype_code(dt), Cc.data_ptr(), st.bufs[k].data_ptr(), n,
                                                     st.flags[k].data_ptr(), st.world, st.steps[k],
                                                     st.err.data_ptr(), T._lib.stream_ptr()))
            e[6].record()
            st.turn ^= 1
        elif pl._transport == "p2p":
            pl.gemm.run(Aa, Bb, st.bufs[st.turn])
            e[4].record()
            k = st.turn
            st.hdls[k].barrier(channel=0)
            e[5].record()
            st.turn = k
            # pull kernel only (gather_p2p would issue a second barrier): re-use its argument builder
            import ctypes as C
            segs = [(st.peer_ptrs[k][r], lo, hi) for r, (lo, hi) in enumerate(pl.ranges) if hi > lo]
            n = len(segs)
            T._lib.check(T._lib.load().t10_gather_peers(T._lib.dtype_code(dt), Cc.data_ptr(), n,
                         (C.c_void_p * n)(*[p for p, _, _ in segs]), (C.c_int64 * n)(*[x for _, x, _ in segs]),
                         (C.c_int64 * n)(*[y for _, _, y in segs]), T._lib.stream_ptr()))
            e[6].record()
            st.turn ^= 1
        else:
            pl.gemm.run(Aa, Bb, Cc)
            e[4].record()
            e[5].record()
            if pl._transport == "nccl":
                parallel.gather_ranges(Cc, pl.ranges, pl.shard[0])
            e[6].record()
        torch.cuda.synchronize()
        for i, k_ in enumerate(names[:6]):
            acc[k_].append(e[i].elapsed_time(e[i + 1]) * 1e3)
        acc["total"].append(e[0].elapsed_time(e[6]) * 1e3)
    out = {k_: float(np.median(v)) for k_, v in acc.items()}
    # max over ranks of the total
    t = torch.tensor([out["total"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["total_max_over_ranks"] = float(t.item())
    out["tiles"] = pl.gemm.ntiles
    out["share_of_flops"] = pl.gemm.flops / pl.flops_total
    parallel.disable()
    return out


for tr in os.environ.get("TRANSPORTS", "fused,p2p,nccl,none").split(","):
    r = run(tr)
    if rank == 0:
        print(json.dumps({"transport": tr, "world": world, **{k: round(v, 2) if isinstance(v, float) else v for k, v in r.items()}}), flush=True)
dist.destroy_process_group()
