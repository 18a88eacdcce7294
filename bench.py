#!/usr/bin/env python
"""bench.py -- U1 block-sparse Contract GFLOP/s (fp64) at chi=4096 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chi 4096]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `Contract(T1, T2)` of workload C3a: legs v = gauss_u1(chi,-20,19,6.0) and
p = {+1,-1}; T1[(a,s);(t,b)] stored as [a,t,s,b] (labels [0,2,1,3], needs the block relayout), T2[(t,b);(u,c)]
(labels [2,3,4,5]); 42 matched sectors at chi=4096, 2.74 GFLOP algorithmic (sum_q 2 m_q k_q n_q).
The step runs: family-2 block relayout of T1 -> family-3 grouped DMMA GEMM (plan cached, as in any loop
that contracts the same bond structure repeatedly, e.g. iTEBD / DMRG sweeps).

JSON line keys follow the driver contract: value (device-timed, inputs resident in HBM), e2e (host buffers,
H2D + Contract + D2H), roofline (grouped GEMM vs FP64 peak measured live), roofline_relayout (HBM),
cpu_baseline (oracle port on the host cores, bounded sample), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NCU_RAW = os.path.join(ROOT, "profiles", "r02_ncu_raw_final.csv")     # `ncu --set full` of this command, raw page


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, read from the COMMITTED ncu capture of this
    command (profiles/r02_ncu_raw_final.csv, C3a chi=4096, 1 GPU).  Fails loudly when the capture holds no launch of
    the kernel the roofline names: a stale capture must not lend its traffic to another kernel."""
    import csv
    if not os.path.exists(NCU_RAW):
        return None
    rows = list(csv.reader(open(NCU_RAW)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in rows[2:] if kernel in r[ki]]
    assert vals, "profiles/r02_ncu_raw_final.csv holds no launch of %s: re-capture it" % kernel
    return float(np.mean(vals))
METRIC = "U1 block-sparse Contract GFLOP/s (fp64) at chi=4096"      # (+ "; iTEBD steps/sec": the `itebd` object)
UNIT = "GFLOP/s"


def c3a_specs(chi):
    from oracle import tor10_oracle as orc  # generator of the synthetic legs only (no arithmetic)
    v = orc.gauss_u1_qnums(chi, -20, 19, 6.0).tolist()
    p = [[1], [-1]]
    u1 = [["U1", 0]]
    A = {"bonds": [{"dim": chi, "btype": -1, "qnums": v, "syms": u1}, {"dim": 2, "btype": 1, "qnums": p, "syms": u1},
                   {"dim": 2, "btype": -1, "qnums": p, "syms": u1}, {"dim": chi, "btype": 1, "qnums": v, "syms": u1}],
         "rowrank": 2, "labels": [0, 2, 1, 3]}
    B = {"bonds": [{"dim": 2, "btype": -1, "qnums": p, "syms": u1}, {"dim": chi, "btype": -1, "qnums": v, "syms": u1},
                   {"dim": 2, "btype": 1, "qnums": p, "syms": u1}, {"dim": chi, "btype": 1, "qnums": v, "syms": u1}],
         "rowrank": 2, "labels": [2, 3, 4, 5]}
    return A, B


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def cpu_port_run(chi, budget_s, max_reps):
    """The oracle (vectorised numpy relayout + torch.matmul per sector, all host threads) on workload C3a."""
    import torch
    from helpers import osym
    from oracle import tor10_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    A, B = c3a_specs(chi)
    oa, ob = osym(A, seed=1003), osym(B, seed=1004)
    mm = lambda x, y: torch.matmul(torch.from_numpy(x), torch.from_numpy(y)).numpy()  # noqa: E731
    flops = None
    times = []
    t_end = time.time() + budget_s
    while len(times) < max_reps and (time.time() < t_end or len(times) < 2):
        t0 = time.perf_counter()
        oc = orc.contract_sym(oa, ob, matmul=mm)
        times.append(time.perf_counter() - t0)
        if flops is None:
            # out sector q pairs with A' and B' sectors of equal charge; A' is [m_q, k_q], out is [m_q, n_q]
            aprime = oa.clone().permute([0, 2, 1, 3], rowrank=2)
            from oracle.tor10_oracle import build_blockmap
            bm = build_blockmap(aprime.bonds, aprime.braket, 2)
            kq = {tuple(q): s[1] for q, s in zip(bm.block_qnums.tolist(), bm.shapes)}
            flops = sum(2.0 * blk.shape[0] * blk.shape[1] * kq[tuple(q)]
                        for q, blk in zip(oc.bm.block_qnums.tolist(), oc.blocks) if tuple(q) in kq)
    return flops, times


def oracle_parity(chi, c):
    """Checker (not timed, not shipped): the oracle's Contract of the same seeded operands against the product's
    result `c` -- block map bit-exact, every block within max|d| / max|ref| (fp64 bar: 1e-12)."""
    import torch
    from helpers import osym, rel_err
    from oracle import tor10_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    A, B = c3a_specs(chi)
    mm = lambda x, y: torch.matmul(torch.from_numpy(x), torch.from_numpy(y)).numpy()  # noqa: E731
    oc = orc.contract_sym(osym(A, seed=1003), osym(B, seed=1004), matmul=mm)
    maps_equal = bool(np.array_equal(c._block_qnums, oc.bm.block_qnums)
                      and np.array_equal(c._Ket_mapper_blks, oc.bm.ket_map)
                      and np.array_equal(c._Bra_mapper_blks, oc.bm.bra_map)
                      and [tuple(x.shape) for x in c.Storage] == [tuple(x.shape) for x in oc.blocks])
    worst = 0.0
    if maps_equal:
        for blk, ref in zip(c.Storage, oc.blocks):
            worst = max(worst, rel_err(blk.cpu().numpy(), ref))
    return {"maps_equal": maps_equal, "max_rel_err": worst if maps_equal else None, "tolerance": 1e-12,
            "blocks": len(oc.blocks), "vs": "oracle.contract_sym (numpy restatement of UniTensor.py:2108-2129, "
            ":3692-3743, pinned to the reference's fixtures) on the same seeded operands, full size"}


def itebd_leg(budget_s=25.0):
    """Second half of the BASELINE metric: iTEBD steps per second (config 1: `iTEBD.py 1.0 0.5 32 1e-8`, untagged
    tensors, fp64, the whole run to convergence from a seeded random state).  `steps_per_s` is the B200-native path
    (the step captured into one CUDA graph, examples/itebd.py run_graphed); `eager_steps_per_s` is the unmodified
    drop-in script path (one Python call per library call, `.item()` every step); `cpu_port_steps_per_s` is the
    numpy restatement of the reference loop on the host cores (checker / baseline only)."""
    import importlib.util
    import torch
    from oracle import itebd_numpy as it
    spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    chi, J, Hx, cvg = 32, 1.0, 0.5, 1e-8
    rng = np.random.Generator(np.random.PCG64(1234))
    init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
    cap = 20000
    ex.run(J, Hx, chi, 0.0, max_steps=20, init=init)                      # plans / caches warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rg = ex.run_graphed(J, Hx, chi, cvg, max_steps=cap, init=init)
    torch.cuda.synchronize()
    tg = time.perf_counter() - t0
    n_eager = min(rg["steps"], max(200, int(budget_s * 0.4 / 0.002)))    # a bounded prefix of the same run
    t0 = time.perf_counter()
    re_ = ex.run(J, Hx, chi, 0.0, max_steps=n_eager, init=init)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    t0 = time.perf_counter()
    rc = it.run(J, Hx, chi, cvg, max_steps=cap, init=init)
    tc = time.perf_counter() - t0
    ne_ = min(len(re_["energies"]), len(rc["energies"]))
    ng = min(len(rg["energies"]), len(rc["energies"]))
    return {"config": "transverse-field Ising chain J=1.0 Hx=0.5 chi=32 cvg=1e-8, untagged UniTensors, fp64, PCG64(1234) start",
            "steps": rg["steps"], "E": rg["E"], "steps_per_s": rg["steps"] / tg, "ms_per_step": tg / rg["steps"] * 1e3,
            "path": "step captured into one CUDA graph (tor10_b200.graph.StaticLoop); energy read back every step",
            "eager_steps_per_s": n_eager / te, "eager_ms_per_step": te / n_eager * 1e3, "eager_steps_timed": n_eager,
            "cpu_port_steps_per_s": rc["steps"] / tc, "cpu_port_steps": rc["steps"], "E_cpu_port": rc["E"],
            "cores": os.cpu_count() or 1,
            "max_abs_dE_vs_cpu_port": float(np.abs(np.array(rg["energies"][:ng]) - np.array(rc["energies"][:ng])).max()),
            "max_abs_dE_eager_vs_cpu_port": float(np.abs(np.array(re_["energies"][:ne_]) - np.array(rc["energies"][:ne_])).max()),
            "svd": "one-CTA Jacobi kernel (t10_svd_small, warm-started); cuSOLVER gesvdj on this matrix: 1.8 ms"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-itebd", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="warmup + steps only (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "C3a: U1 rank-4 x rank-4 blockform Contract, chi=%d, d=2, gauss_u1(-20..19, sigma 6), fp64" % args.chi

    # ------------------------------------------------------------------ reference arm -----
    if args.impl == "reference":
        if rank != 0:
            return
        flops, times = cpu_port_run(args.chi, budget_s=max(20.0, 0.6 * (args.steps + args.warmup)),
                                    max_reps=args.steps + args.warmup)
        timed = times[min(args.warmup, len(times) - 1):] or times
        ms = 1e3 * float(np.mean(timed))
        val = flops / (ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "note": "reference CPU path restated (oracle port: numpy relayout + "
                           "torch.matmul per sector); the unmodified Python reference cannot travel to the GPU box "
                           "and needs ~60-80 s per operand relayout at this size (SURVEY 6)"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d full Contract calls of the workload" % len(timed)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm -----------
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            # NCCL logs to stdout by default and the bench prints ONE JSON line there: keep the level the caller
            # asked for and send the log (communicator / rank lines included) to stderr instead
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    import tor10_b200 as T
    from helpers import psym
    from tor10_b200 import _plan, parallel

    default_transport = os.environ.get("TOR10_GATHER", "auto")
    if world > 1:
        parallel.enable(transport=default_transport)
    A, B = c3a_specs(args.chi)
    a = psym(T, A, seed=1003, device=dev)
    b = psym(T, B, seed=1004, device=dev)
    plan = _plan.get_contract_plan(a, b)
    flops = plan.flops_total                       # algorithmic, whole job
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, n, pre=None):
        """n calls, each bracketed by its own CUDA events; L2 flushed (256 MiB write) between calls."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for s, e in evs:
            if pre is not None:
                pre()
            flush.fill_(1.0)
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in evs]

    out_holder = [None]

    def step():
        out_holder[0] = T.Contract(a, b)

    for _ in range(args.warmup):
        step()
    barrier()
    if args.profile_only:
        for _ in range(args.steps):
            flush.fill_(1.0)
            step()
        torch.cuda.synchronize()
        return
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.perf_counter()
    per_step = timed_loop(step, args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_local = float(np.sum(per_step))
    if world > 1:
        tt = torch.tensor([ms_local], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    else:
        ms_total = ms_local
    ms_per_step = ms_total / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e9

    # ---- the same step through the other transports: "pull" / "fused" gather the result on every rank --------
    others = {}
    if world > 1:
        for tr in ("resident", "push", "pull", "fused"):
            if tr == plan._transport:
                continue
            try:
                parallel.enable(transport=tr)
                for _ in range(3):
                    step()
                barrier()
                ts = float(np.sum(timed_loop(step, args.steps)))
                tt = torch.tensor([ts], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item()) / args.steps
                step()
                c_tr = out_holder[0]
                c_tr._materialize()
                torch.cuda.synchronize()
                barrier()
                others[tr] = {"ms_per_step": ms, "value": flops / (ms * 1e-3) / 1e9,
                              "parity": oracle_parity(args.chi, c_tr) if rank == 0 else None}
                barrier()
            except Exception as exc:   # a transport that cannot run here is reported, not hidden
                others[tr] = {"error": repr(exc)[:300]}
        out_holder[0] = None
        parallel.enable(transport=default_transport)
        plan = _plan.get_contract_plan(a, b)

    # ---- stage timings on this rank's share (roofline of the dominant kernel) ---------------
    A_arena, B_arena = a._arena_tensor(), b._arena_tensor()
    dt = a.dtype
    Ap = torch.empty(plan.rel_a.packed.size, device=dev, dtype=dt) if plan.rel_a is not None else A_arena
    Bp = torch.empty(plan.rel_b.packed.size, device=dev, dtype=dt) if plan.rel_b is not None else B_arena
    Cc = plan.Lo.new_arena(dt, zero=True)
    if plan.rel_a is not None:
        plan.rel_a.run(A_arena, Ap)
    if plan.rel_b is not None:
        plan.rel_b.run(B_arena, Bp)
    nrep = max(10, min(args.steps, 50))
    t_gemm = timed_loop(lambda: plan.gemm.run(Ap, Bp, Cc), nrep)
    t_rel = timed_loop(lambda: plan.rel_a.run(A_arena, Ap), nrep) if plan.rel_a is not None else [0.0]
    gemm_ms, rel_ms = float(np.mean(t_gemm)), float(np.mean(t_rel))
    rel_bytes = 2.0 * 8.0 * plan.rel_a.moved_elems if plan.rel_a is not None else 0.0

    # ---- end to end through the public API with HOST buffers -----------------------------------
    hA = A_arena.cpu().pin_memory()
    hB = B_arena.cpu().pin_memory()
    hC = torch.empty(plan.Lo.arena_elems + 16, dtype=dt).pin_memory()

    def merged(ranges):
        out = []
        for lo, hi in sorted(ranges):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(int(lo), int(hi)) for lo, hi in out]

    def source_ranges(rel, layout, direct_sectors):
        """Arena ranges of the operand this rank's share of the contraction reads."""
        if rel is None:
            secs = direct_sectors
        elif rel.d_seg is not None:
            sg = rel.d_seg.cpu().numpy().astype(np.int64)
            sg = sg[sg[:, 1] >= 0]
            first, last = sg[:, 1], sg[:, 1] + sg[:, 2] - 1
            secs = np.unique(np.concatenate([np.searchsorted(layout.arena_off, first, side="right") - 1,
                                             np.searchsorted(layout.arena_off, last, side="right") - 1]))
        elif rel.d_runs is not None:
            rr = rel.d_runs.cpu().numpy().astype(np.int64)
            ln = np.diff(rr[:, 0])
            ok = rr[:-1, 1] >= 0
            first, last = rr[:-1, 1][ok], rr[:-1, 1][ok] + ln[ok] - 1
            secs = np.unique(np.concatenate([np.searchsorted(layout.arena_off, first, side="right") - 1,
                                             np.searchsorted(layout.arena_off, last, side="right") - 1]))
        elif rel.d_index is None:
            return [(0, int(layout.arena_elems))]
        else:
            idx = rel.d_index.cpu().numpy()
            idx = idx[idx >= 0]
            secs = np.unique(np.searchsorted(layout.arena_off, idx, side="right") - 1)
        return merged([(layout.arena_off[s_], layout.arena_off[s_] + layout.rows[s_] * layout.ld[s_]) for s_ in secs])

    if world > 1:
        # every rank moves only what its sectors need: the source sectors its relayouts read, and its own output
        # sectors (the union over the ranks is each operand / the result exactly once per step)
        in_a = source_ranges(plan.rel_a, a._layout, sorted(set(m[1] for m in plan.matched)))
        in_b = source_ranges(plan.rel_b, b._layout, sorted(set(m[2] for m in plan.matched)))
        out_c = merged([(plan.Lo.arena_off[m[0]], plan.Lo.arena_off[m[0]] + plan.Lo.rows[m[0]] * plan.Lo.ld[m[0]])
                        for m in plan.matched])
    else:
        in_a, in_b = [(0, hA.numel())], [(0, hB.numel())]
        out_c = [(0, hC.numel())]
    h2d_bytes = 8 * (sum(hi - lo for lo, hi in in_a) + sum(hi - lo for lo, hi in in_b))
    d2h_bytes = 8 * sum(hi - lo for lo, hi in out_c)

    def e2e_step():
        for lo, hi in in_a:
            a._arena[lo:hi].copy_(hA[lo:hi], non_blocking=True)
        for lo, hi in in_b:
            b._arena[lo:hi].copy_(hB[lo:hi], non_blocking=True)
        c = T.Contract(a, b)
        for lo, hi in out_c:
            hC[lo:hi].copy_(c._arena[lo:hi], non_blocking=True)     # (this rank's own sectors when sharded)

    for _ in range(3):
        e2e_step()
    barrier()
    t_e2e = timed_loop(e2e_step, max(5, min(args.steps, 20)))
    barrier()
    e2e_serial_ms = float(np.mean(t_e2e))

    # The same steps as a stream of work: upload of step i+1, Contract of step i and download of step i-1
    # run concurrently on three CUDA streams (two device operand sets, two pinned result buffers).  Every
    # step still uploads its own operands and downloads its own result inside the timed region; PCIe is
    # full duplex, so the step time drops from H2D + Contract + D2H to about the upload time.
    import copy as _copy
    sets = [(a, b), (_copy.deepcopy(a), _copy.deepcopy(b))]
    hCs = [hC, torch.empty_like(hC).pin_memory()]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def e2e_stream(nsteps):
        cur = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        keep = [None, None]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        t0.record(cur)
        for i in range(nsteps):
            k = i & 1
            ta, tb = sets[k]
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_done[k])          # the Contract of step i-2 has read this operand set
                else:
                    s_in.wait_event(t0)
                for lo, hi in in_a:
                    ta._arena[lo:hi].copy_(hA[lo:hi], non_blocking=True)
                for lo, hi in in_b:
                    tb._arena[lo:hi].copy_(hB[lo:hi], non_blocking=True)
                ev_in[k].record(s_in)
            cur.wait_event(ev_in[k])
            c = T.Contract(ta, tb)
            ev_done[k].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                for lo, hi in out_c:
                    hCs[k][lo:hi].copy_(c._arena[lo:hi], non_blocking=True)
                c._arena.record_stream(s_out)
            keep[k] = c
        cur.wait_stream(s_out)
        t1.record(cur)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / nsteps

    n_e2e = max(10, min(args.steps, 40))
    e2e_stream(4)
    barrier()
    e2e_ms = e2e_stream(n_e2e)
    barrier()
    if world > 1:
        tt = torch.tensor([e2e_ms, e2e_serial_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_serial_ms = float(tt[0].item()), float(tt[1].item())
    del sets, hCs
    # ---- the same contraction in fp32 (tcgen05 / TMEM 3xTF32 grouped GEMM), informational, 1 GPU only ----
    fp32_info = None
    if world == 1:
        a32 = psym(T, A, seed=1003, device=dev, dtype=torch.float32)
        b32 = psym(T, B, seed=1004, device=dev, dtype=torch.float32)
        for _ in range(3):
            T.Contract(a32, b32)
        t32 = float(np.mean(timed_loop(lambda: T.Contract(a32, b32), max(10, min(args.steps, 50)))))
        p32 = _plan.get_contract_plan(a32, b32)
        fp32_info = {"ms_per_step": t32, "gflops": flops / (t32 * 1e-3) / 1e9, "tile_cfg": p32.gemm.tile_cfg,
                     "kernel": "k_ggemm_tf32x3 (tcgen05.mma kind::tf32, 3-pass split, TMEM accumulators)"
                     if p32.gemm.tile_cfg == 50 else "k_ggemm_simt<float>"}
        del a32, b32
    clocks = sampler.stop() if sampler is not None else None

    # ---- parity of the headline result against the oracle (checker, outside every timed region) ----------
    step()
    if world > 1:
        out_holder[0]._materialize()      # sector-resident result: gathered on demand, on every rank
    torch.cuda.synchronize()
    barrier()
    parity = oracle_parity(args.chi, out_holder[0]) if rank == 0 else None
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- FP64 peak, measured live (MEASURED_PEAKS.json carries no fp64 figure) -------------------
    import ctypes
    lib = T._lib.load()
    tf = ctypes.c_double(0.0)
    # peak policy: three repetitions of every probe, the best of each is kept, all are printed; the denominator
    # is never below the cuBLAS DGEMM figure
    dmma_runs, dfma_runs, dgemm_runs = [], [], []
    n = 8192
    x = torch.rand((n, n), device=dev, dtype=torch.float64)
    y = torch.rand((n, n), device=dev, dtype=torch.float64)
    torch.matmul(x, y)
    for _ in range(3):
        T._lib.check(lib.t10_probe_f64_peak(1, 4096, ctypes.byref(tf), T._lib.stream_ptr()))
        dmma_runs.append(tf.value)
        T._lib.check(lib.t10_probe_f64_peak(0, 4096, ctypes.byref(tf), T._lib.stream_ptr()))
        dfma_runs.append(tf.value)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(x, y)
        e.record()
        torch.cuda.synchronize()
        dgemm_runs.append(2.0 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12)
    del x, y
    dmma_probe, dfma_probe, dgemm_tf = max(dmma_runs), max(dfma_runs), max(dgemm_runs)
    peak_tf = max(dgemm_tf, dmma_probe)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    share = plan.gemm.flops / flops if flops else 1.0
    gemm_tf = plan.gemm.flops / (gemm_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sectors": int(plan.Lo.nsec), "algorithmic_gflop": flops / 1e9,
                   "nnz_per_operand": int(a._layout.nnz), "tile_cfg": plan.gemm.tile_cfg,
                   "parallelism": ("sectors sharded over %d GPUs, LPT by flops; transport %s: %s" % (
                       world, plan._transport,
                       {"resident": "results stay sector-resident (every rank keeps the sectors it computed, owner table "
                                    "on the tensor, later contractions read them in place or fetch foreign sectors over "
                                    "NVLink peer memory, gathered on demand); no collective on the data path",
                        "pull": "resident, then every rank pulls the other ranks' sectors over NVLink peer memory",
                        "push": "all-gather fused into the TMA-fed GEMM: finished tiles stored from the accumulator "
                                "registers into every rank's ring slot over NVLink, progress words, no staging copy",
                        "fused": "all-gather fused into the GEMM epilogue (tiles stored into all ranks' symmetric "
                                 "staging arenas over NVLink, system-scope flags, one wait+copy kernel)",
                        "p2p": "symmetric-memory staging + device barrier + one pull kernel",
                        "nccl": "NCCL all_gather_into_tensor"}.get(plan._transport, str(plan._transport))))
                   if world > 1 else "1 GPU",
                   "other_transports": others if world > 1 else None,
                   "l2": "flushed between steps (256 MiB write); every step timed by its own CUDA event pair",
                   "plan": "cached (block maps, relayout tables, tile table built once outside the timed region)"},
        "e2e": {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "serial_ms_per_step": e2e_serial_ms, "serial_value": flops / (e2e_serial_ms * 1e-3) / 1e9,
                "note": ("steps streamed over three CUDA streams (upload of step i+1 | Contract of step i | download "
                         "of step i-1, two operand sets, working set > L2); serial_* = one step at a time (upload, "
                         "Contract, download back to back, L2 flushed). "
                         + ("Bytes are rank 0's: every rank uploads the operand sectors its share reads and "
                            "downloads its own output sectors" if world > 1 else "Whole operands in, whole result out"))},
        "fp32_contract": fp32_info,
        "gpu_launches": int(args.steps * ((1 if plan.rel_a is not None else 0) + (1 if plan.rel_b is not None else 0)
                                          + 1 + ((1 if plan._transport in ("fused", "p2p", "pull", "push") else 0) if world > 1 else 0))),
        "roofline": {"kernel": ("k_ggemm_f64_tma (grouped DMMA GEMM, TMA producer warp + mbarrier ring, mixed tile "
                                "shapes, stream-K)" if getattr(plan.gemm, "tma", False) else
                                "k_ggemm_f64_mixed_sk (grouped DMMA GEMM, mixed tile shapes, stream-K)"
                                if getattr(plan.gemm, "streamk", False) else
                                "k_ggemm_f64_mixed (grouped DMMA GEMM, mixed tile shapes)" if plan.gemm.tile_cfg == 46 else
                                "k_ggemm_f64_v5 (grouped DMMA GEMM, aligned software-pipelined kernel)"
                                if plan.gemm.tile_cfg in (40, 42) else "k_ggemm_f64 (grouped DMMA GEMM)"),
                     "bound": "tensor", "achieved": gemm_tf,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf if peak_tf else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
                     # command at chi=4096 on 1 GPU (profiles/r01_ncu_summary.md): 49.97 MB + 4.21 MB (tile_cfg 40),
                     # 50.00 MB + 3.02 MB (tile_cfg 3)
                     # 50.74 MB + 11.09 MB (tile_cfg 46 as stream-K: the parked partial tiles add ~8 MB)
                     "traffic": (ncu_traffic("k_ggemm_f64_tma") if getattr(plan.gemm, "tma", False) else None)
                     if (world == 1 and args.chi == 4096) else None,
                     "traffic_source": "profiles/r02_ncu_raw_final.csv (ncu --set full of `bench.py --steps 2 --warmup 3 "
                                       "--profile-only`, dram read + write per launch, L2 flushed before the launch)",
                     "algorithmic_bytes": 8.0 * plan.gemm.bytes, "ms": gemm_ms, "rank_share_of_flops": share,
                     "peak_source": "FP64 measured live in this run, best of 3 each: max(torch.matmul f64 8192^3 = %.1f TF, "
                                    "DMMA register probe = %.1f TF; DFMA probe = %.1f TF); MEASURED_PEAKS.json has "
                                    "no fp64 entry" % (dgemm_tf, dmma_probe, dfma_probe),
                     "peak_runs": {"dmma_probe": dmma_runs, "dfma_probe": dfma_runs, "dgemm_8192": dgemm_runs}},
        "roofline_relayout": {"kernel": ("k_relayout_segments (block relayout replayed through the contiguous segments "
                                         "of its element index, one warp per segment)"
                                         if (plan.rel_a is not None and plan.rel_a.d_seg is not None) else
                                         "k_relayout_runs (block relayout replayed through the runs of its element index)"
                                         if (plan.rel_a is not None and plan.rel_a.d_runs is not None) else
                                         "k_relayout_index (block relayout replayed through its element index)"
                                         if (plan.rel_a is not None and plan.rel_a.d_index is not None) else
                                         "k_relayout_blocks (factored tables)"), "bound": "hbm",
                              "achieved": rel_bytes / (rel_ms * 1e-3) / 1e9 if rel_ms else None, "peak": hbm_peak,
                              "unit": "GB/s", "frac": rel_bytes / (rel_ms * 1e-3) / 1e9 / hbm_peak if rel_ms else None,
                              "ms": rel_ms, "algorithmic_bytes": rel_bytes, "peak_source": hbm_src,
                              # dram read+write of one launch, ncu --set full at chi=4096 on 1 GPU
                              # (profiles/r01_ncu_summary.md); the destination partly stays in L2
                              "traffic": (ncu_traffic("k_relayout_segments")
                                          if (plan.rel_a is not None and plan.rel_a.d_seg is not None) else None)
                              if (world == 1 and args.chi == 4096) else None},
        "parity": parity,
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if world == 1 and not args.no_itebd:
        try:
            line["itebd"] = itebd_leg()
        except Exception as exc:
            line["itebd"] = {"error": repr(exc)[:300]}
    if world == 1 and not args.no_cpu_baseline:
        cf, ctimes = cpu_port_run(args.chi, budget_s=12.0, max_reps=20)
        ct = float(np.mean(ctimes[1:] or ctimes))
        line["cpu_baseline"] = {"value": cf / ct / 1e9, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "%d full Contract calls of the same workload by the oracle (numpy "
                                          "vectorised relayout + torch.matmul per sector)" % len(ctimes[1:] or ctimes)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
