#!/usr/bin/env python
"""Secondary measurements of SURVEY 8(d) (not the driver's bench line): prints one JSON object per item.

  C1   iTEBD chi=32 (examples/itebd.py): steps/s on the GPU path, Svd_truncate share timed separately,
       numpy port of the reference loop on the host cores beside it
  C2   dense rank-4 Contract chi=1024, d=2 (permute + GEMM): GFLOP/s, permute GB/s alone
  C3a  stage split + the reference torch path ON THE GPU (per-sector torch.matmul loop on pre-permuted
       blocks = the cuBLAS bar), C3b MPS-MPO-MPS environment step
  F1   block-map build latency (family 1) at C3a sizes, plan build (cold) vs replay (warm)
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T  # noqa: E402
from helpers import psym  # noqa: E402
from oracle import tor10_oracle as orc  # noqa: E402  (synthetic-leg generator + host baselines only)

RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if WORLD > 1:       # torchrun: c4 / c5 with sectors sharded over the ranks, results sector-resident (tor10_b200.parallel)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    from tor10_b200 import parallel
    parallel.enable(transport=os.environ.get("TOR10_GATHER", "resident"))
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)


def _max_over_ranks(x):
    if WORLD == 1:
        return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, n=30, warm=3):
    for _ in range(warm):
        fn()
    if WORLD > 1:
        torch.cuda.synchronize()
        dist.barrier()
    evs = []
    for _ in range(n):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return _max_over_ranks(float(np.median([s.elapsed_time(e) for s, e in evs])))


def emit(**kw):
    if RANK == 0:
        print(json.dumps(dict(kw, n_gpus=WORLD)), flush=True)


def c1_itebd(chi=32, steps=300):
    spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    from oracle import itebd_numpy as it
    rng = np.random.Generator(np.random.PCG64(7))
    init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
    ex.run(1.0, 0.5, chi, 0.0, max_steps=20, init=init, check_every=10 ** 9)      # warm-up: plans, cuSOLVER handles
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ex.run(1.0, 0.5, chi, 0.0, max_steps=steps, init=init, check_every=steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # Svd_truncate alone (torch/cuSOLVER on device), same matrix size
    M = T.UniTensor(bonds=[T.Bond(2 * chi), T.Bond(2 * chi)], rowrank=1, device=dev).Rand()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(50):
        T.Svd_truncate(M, chi)
    torch.cuda.synchronize()
    svd_ms = (time.perf_counter() - t1) / 50 * 1e3
    t2 = time.perf_counter()
    ref = it.run(1.0, 0.5, chi, 0.0, max_steps=steps, init=init)
    dtc = time.perf_counter() - t2
    emit(item="C1 iTEBD TFIM chi=%d" % chi, steps=steps, steps_per_s=steps / dt, ms_per_step=dt / steps * 1e3,
         svd_truncate_ms=svd_ms, E=out["E"], E_cpu_port=ref["E"], abs_dE=abs(out["E"] - ref["E"]),
         cpu_port_steps_per_s=steps / dtc, cpu_cores=os.cpu_count(),
         note="wall clock incl. Python host code; one .item() sync at the end; reference CPU (8 vCPU, build "
              "container): 218 steps/s")


def c2_dense(chi=1024, d=2):
    g = torch.Generator(device="cpu").manual_seed(2)
    t1 = T.UniTensor(bonds=[T.Bond(chi), T.Bond(d), T.Bond(d), T.Bond(chi)], rowrank=2, labels=[0, 2, 1, 3], device=dev)
    t2 = T.UniTensor(bonds=[T.Bond(d), T.Bond(chi), T.Bond(d), T.Bond(chi)], rowrank=2, labels=[2, 3, 4, 5], device=dev)
    t1.Storage = torch.rand(t1.Storage.shape, generator=g, dtype=torch.float64).to(dev)
    t2.Storage = torch.rand(t2.Storage.shape, generator=g, dtype=torch.float64).to(dev)
    flops = 2.0 * (chi * d) ** 3
    ms = timed(lambda: T.Contract(t1, t2), n=20)
    want = torch.tensordot(t1.Storage, t2.Storage, dims=([1, 3], [0, 1]))
    got = T.Contract(t1, t2).Storage
    err = float((got - want).abs().max() / want.abs().max())
    ms_torch = timed(lambda: torch.tensordot(t1.Storage, t2.Storage, dims=([1, 3], [0, 1])), n=20)
    x = t1.Storage.permute(0, 2, 1, 3)
    ms_perm = timed(lambda: T.UniTensor._contiguous_dense(x), n=30)
    y = t1.Storage.reshape(chi * d, d * chi).t()
    ms_tr = timed(lambda: T.UniTensor._contiguous_dense(y), n=30)
    nbytes = 2.0 * 8 * t1.Storage.numel()
    emit(item="C2 dense rank-4 Contract chi=%d d=%d" % (chi, d), gflops=flops / ms / 1e6, ms=ms, rel_err_vs_torch=err,
         torch_tensordot_gpu_gflops=flops / ms_torch / 1e6, permute_rows_GBs=nbytes / ms_perm / 1e6,
         permute_rows_ms=ms_perm, transpose_2048_GBs=nbytes / ms_tr / 1e6, transpose_ms=ms_tr)


def c3(chi=4096):
    from bench import c3a_specs
    A, B = c3a_specs(chi)
    a, b = psym(T, A, seed=1003, device=dev), psym(T, B, seed=1004, device=dev)
    plan = T._plan.get_contract_plan(a, b)
    ms = timed(lambda: T.Contract(a, b), n=50)
    # the reference's torch path on the GPU: per-sector torch.matmul on pre-permuted blocks
    ap = a.Permute([0, 2, 1, 3], rowrank=2).Contiguous() if False else None
    a2 = psym(T, A, seed=1003, device=dev)
    a2.Permute([0, 2, 1, 3], rowrank=2)
    a2c = a2.Contiguous()
    Ab, Bb = list(a2c.Storage), list(b.Storage)
    pairs = [(ab, bb) for _, ab, bb in plan.matched]

    def ref_loop():
        return [torch.matmul(Ab[i], Bb[j]) for i, j in pairs]
    ms_ref = timed(ref_loop, n=30)
    emit(item="C3a U1 Contract chi=%d" % chi, ms=ms, gflops=plan.flops_total / ms / 1e6,
         reference_torch_path_on_gpu_ms=ms_ref, reference_torch_path_on_gpu_gflops=plan.flops_total / ms_ref / 1e6,
         note="reference path = its per-sector torch.matmul loop (UniTensor.py:3727-3741) on already-relayouted "
              "blocks, i.e. cuBLAS DGEMM x%d launches, relayout NOT included" % len(pairs))
    # C3b: L[(x, w); (a)] x A[(a); (s, b)]
    v = orc.gauss_u1_qnums(chi, -20, 19, 6.0).tolist()
    u1 = [["U1", 0]]
    w = [[0], [0], [0], [2], [-2]]
    p = [[1], [-1]]
    Ls = {"bonds": [{"dim": chi, "btype": -1, "qnums": v, "syms": u1}, {"dim": 5, "btype": -1, "qnums": w, "syms": u1},
                    {"dim": chi, "btype": 1, "qnums": v, "syms": u1}], "rowrank": 2, "labels": [10, 11, 12]}
    As = {"bonds": [{"dim": chi, "btype": -1, "qnums": v, "syms": u1}, {"dim": 2, "btype": 1, "qnums": p, "syms": u1},
                    {"dim": chi, "btype": 1, "qnums": v, "syms": u1}], "rowrank": 1, "labels": [12, 13, 14]}
    for sp in (Ls, As):
        for bd in sp["bonds"]:
            bd["qnums"] = sorted(bd["qnums"], reverse=True)
    L, Am = psym(T, Ls, seed=1005, device=dev), psym(T, As, seed=1006, device=dev)
    pl = T._plan.get_contract_plan(L, Am)
    msb = timed(lambda: T.Contract(L, Am), n=50)
    emit(item="C3b MPS-MPO-MPS environment step chi=%d" % chi, sectors=len(pl.matched), gflop=pl.flops_total / 1e9,
         ms=msb, gflops=pl.flops_total / msb / 1e6)


def f1_blockmap(chi=4096):
    from bench import c3a_specs
    A, B = c3a_specs(chi)
    T._plan.clear_caches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a = psym(T, A, device=dev)
    torch.cuda.synchronize()
    t_ctor = time.perf_counter() - t0
    b = psym(T, B, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    T.Contract(a, b)
    torch.cuda.synchronize()
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(20):
        T.Contract(a, b)
    torch.cuda.synchronize()
    t_warm = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        psym(T, A, device=dev)
    torch.cuda.synchronize()
    t_ctor_warm = (time.perf_counter() - t0) / 20
    emit(item="F1 block map / plans chi=%d" % chi, first_UniTensor_ctor_ms=t_ctor * 1e3, cached_ctor_ms=t_ctor_warm * 1e3,
         first_Contract_ms_plan_build=t_cold * 1e3, replay_Contract_wall_ms=t_warm * 1e3,
         note="reference (8 vCPU): block-map build 22.8 ms per tensor, executed 3x per Contract; full Contract "
              "68.3 ms without relayout (SURVEY 6)")


def _u1z2_leg(chi, qlo, qhi, sigma):
    q = orc.gauss_u1_qnums(chi, qlo, qhi, sigma)[:, 0]
    return sorted([[int(x), int(x) % 2] for x in q], reverse=True)


def c4_u1z2(chi=2048):
    """Config 4 hot path: two-site update of a U1xZ2 MPS -- symmetric Contract, then the sector-wise
    Svd_truncate (torch/cuSOLVER on device, timed separately as the north star prescribes)."""
    syms = [["U1", 0], ["Zn", 2]]
    v = _u1z2_leg(chi, -16, 15, 5.0)
    p = [[1, 1], [-1, 1]]

    def bd(dim, bt, qn):
        return {"dim": dim, "btype": bt, "qnums": qn, "syms": syms}
    A = {"bonds": [bd(chi, -1, v), bd(2, -1, p), bd(chi, 1, v)], "rowrank": 2, "labels": [1, 2, 3]}
    B = {"bonds": [bd(2, 1, p), bd(chi, -1, v), bd(chi, 1, v)], "rowrank": 2, "labels": [4, 3, 5]}
    a, b = psym(T, A, seed=1401, device=dev), psym(T, B, seed=1402, device=dev)
    pl = T._plan.get_contract_plan(a, b)
    ms = timed(lambda: T.Contract(a, b), n=30)
    th = T.Contract(a, b)
    T.Svd_truncate(th, chi)                      # (warm: cuSOLVER handles, layouts of the factors)
    th = T.Contract(a, b)
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    t0 = time.perf_counter()
    U, S, V = T.Svd_truncate(th, chi)
    torch.cuda.synchronize()
    svd_s = _max_over_ranks(time.perf_counter() - t0)
    emit(item="C4 U1xZ2 two-site update chi=%d" % chi, sectors=len(pl.matched), gflop=pl.flops_total / 1e9,
         contract_ms=ms, contract_gflops=pl.flops_total / ms / 1e6, svd_truncate_s=svd_s,
         kept=int(S.bonds[0].dim), updates_per_s_contract_only=1e3 / ms, updates_per_s_with_svd=1.0 / (svd_s + ms * 1e-3),
         note="Svd_truncate = %d sector SVDs (torch / cuSOLVER default driver; one-CTA Jacobi kernel for blocks up to "
              "64 wide)%s" % (th._layout.nsec, ", dealt over %d ranks by owner (sector-resident)" % WORLD if WORLD > 1 else ""))


def c5_peps(D=8, chi=512):
    """Config 5: absorbing one double-layer PEPS site into a boundary tensor, as a Network of U1 tensors
    (E: chi x D x D x chi boundary tensor, P and its conjugate Pc: D^4 x d site tensors)."""
    u1 = [["U1", 0]]
    cq = sorted(orc.gauss_u1_qnums(chi, -6, 5, 2.0).tolist(), reverse=True)
    dq = sorted(orc.gauss_u1_qnums(D, -1, 1, 0.8).tolist(), reverse=True)
    pq = [[1], [-1]]

    def bd(dim, bt, qn):
        return {"dim": dim, "btype": bt, "qnums": qn, "syms": u1}
    specs = {
        # E[(x); (u, u', y)]
        "E": {"bonds": [bd(chi, -1, cq), bd(D, 1, dq), bd(D, 1, dq), bd(chi, 1, cq)], "rowrank": 1, "labels": [0, 1, 2, 3]},
        # P[(u, l); (r, d, s)]   Pc[(u', l', s); (r', d')]
        "P": {"bonds": [bd(D, -1, dq), bd(D, -1, dq), bd(D, 1, dq), bd(D, 1, dq), bd(2, 1, pq)], "rowrank": 2,
              "labels": [0, 1, 2, 3, 4]},
        "Pc": {"bonds": [bd(D, -1, dq), bd(D, -1, dq), bd(2, -1, pq), bd(D, 1, dq), bd(D, 1, dq)], "rowrank": 3,
               "labels": [0, 1, 2, 3, 4]},
    }
    net = ("E: 10; 11 12 13\nP: 11 20; 21 22 30\nPc: 12 23 30; 24 25\n"
           "TOUT: 10 20 23; 13 21 22 24 25\nOrder: (E,P),Pc\n")
    N = T.Network()
    N.Fromstring(net)
    ts = {k: psym(T, sp, seed=1500 + i, device=dev) for i, (k, sp) in enumerate(specs.items())}
    for k, t in ts.items():
        N.Put(k, t)
    N.Launch()
    before = dict(T._plan.STATS)
    ms = timed(lambda: N.Launch(), n=20)
    assert T._plan.STATS == before, "relaunch must replay cached plans"
    plans = [pl for key, pl in T._plan._CONTRACTS.items() if isinstance(pl, T._plan.ContractPlan) and pl.Lo.rank >= 5]
    flops = sum(pl.flops_total for pl in plans)
    if WORLD > 1:          # (per-rank byte counts below are rank 0's share)
        torch.cuda.synchronize()
        dist.barrier()
    # algorithmic bytes: GEMM operands + results, plus read+write of every relayouted operand
    nbytes = sum(8.0 * pl.gemm_bytes + sum(16.0 * r.moved_elems for r in (pl.rel_a, pl.rel_b) if r is not None)
                 for pl in plans)
    emit(item="C5 PEPS double-layer site absorption D=%d chi=%d" % (D, chi), ms=ms, gflop=flops / 1e9,
         gflops=flops / ms / 1e6, algorithmic_GB=nbytes / 1e9, GBs=nbytes / ms / 1e6,
         arithmetic_intensity=(flops / nbytes if nbytes else None), relayouts=sum((pl.rel_a is not None) + (pl.rel_b is not None) for pl in plans),
         index_replay=[("segments" if r.d_seg is not None else "runs" if r.d_runs is not None else "index" if r.d_index is not None else "blocks") for pl in plans for r in (pl.rel_a, pl.rel_b) if r is not None],
         note="Network.Launch of 3 U1 tensors (2 Contracts + final Permute), plans replayed; contracting over single "
              "D=8 bonds makes this HBM-bound (K <= 64 per sector)")


if __name__ == "__main__":
    which = sys.argv[1:] or (["c4", "c5", "c2"] if WORLD > 1 else ["f1", "c3", "c2", "c4", "c5", "c1"])
    if "c4" in which:
        c4_u1z2()
    if "c5" in which:
        c5_peps()
    if "f1" in which:
        f1_blockmap()
    if "c3" in which:
        c3()
    if "c2" in which:
        c2_dense()
    if "c1" in which:
        c1_itebd()
        try:
            c1_itebd(chi=64, steps=100)
        except Exception as e:   # large chi: truncated singular values reach 0 and Inverse(lb) fails, as in the reference
            emit(item="C1 iTEBD chi=64", failed=str(e)[:200])
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()
