"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): a Contract whose sectors are sharded over
2 ranks gives bit-identical results through every transport, equals the single-GPU result to rounding (1e-14)
and the oracle within 1e-12."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TOR10_SHARD_MIN_MFLOP="0",
                      TOR10_DENSE_SHARD_MIN_GFLOP="0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import tor10_b200 as T
    from tor10_b200 import parallel
    from helpers import osym, psym, rel_err
    from oracle import tor10_oracle as orc
    from test_gpu_parity import _c3_spec
    A, B = _c3_spec(160, -7, 6, 2.5)
    a, b = psym(T, A, seed=41, device="cuda:%d" % rank), psym(T, B, seed=42, device="cuda:%d" % rank)
    single = T.Contract(a, b)
    a32 = psym(T, A, seed=41, device="cuda:%d" % rank, dtype=torch.float32)
    b32 = psym(T, B, seed=42, device="cuda:%d" % rank, dtype=torch.float32)
    single32 = T.Contract(a32, b32)
    oc = orc.contract_sym(osym(A, seed=41), osym(B, seed=42))
    same, err, first = True, 0.0, None
    # every transport, three calls each (the staging arenas are double-buffered); "auto" = fused GEMM+all-gather.
    # All transports run the same kernels on the same sector shards: their results are bit-identical.  The
    # single-GPU result may come from the stream-K schedule (another, equally fixed, summation split of k), so
    # it is compared at rounding level.
    for transport in ("auto", "fused", "p2p", "nccl"):
        parallel.enable(transport=transport)
        for _ in range(3):
            sharded = T.Contract(a, b)
            if first is None:
                first = [x.clone() for x in sharded.Storage]
            same = same and all(torch.equal(x, y) for x, y in zip(first, sharded.Storage))
            same = same and all(float((x - y).abs().max()) <= 1e-14 * float(x.abs().max())
                                for x, y in zip(single.Storage, sharded.Storage))
            err = max(err, max(rel_err(x.cpu().numpy(), y) for x, y in zip(sharded.Storage, oc.blocks)))
        sh32 = T.Contract(a32, b32)      # fp32 plans cannot fuse: pull path
        same = same and all(torch.equal(x, y) for x, y in zip(single32.Storage, sh32.Storage))
        parallel.disable()
        T._plan.clear_caches()
    # untagged dense tensors: the GEMM is partitioned by output row tiles over the ranks, one NCCL all-gather
    # assembles the result on every rank (config 2 shape at chi = 256)
    chi, d = 256, 2
    g = torch.Generator().manual_seed(3)
    x = torch.rand((chi, d, d, chi), generator=g, dtype=torch.float64).to("cuda:%d" % rank)
    y = torch.rand((d, chi, d, chi), generator=g, dtype=torch.float64).to("cuda:%d" % rank)
    t1 = T.UniTensor(bonds=[T.Bond(chi), T.Bond(d), T.Bond(d), T.Bond(chi)], labels=[0, 2, 1, 3], rowrank=2,
                     device=torch.device("cuda:%d" % rank))
    t2 = T.UniTensor(bonds=[T.Bond(d), T.Bond(chi), T.Bond(d), T.Bond(chi)], labels=[2, 3, 4, 5], rowrank=2,
                     device=torch.device("cuda:%d" % rank))
    t1.Storage, t2.Storage = x, y
    want = torch.tensordot(x, y, dims=([1, 3], [0, 1]))
    one = T.Contract(t1, t2).Storage.clone()
    parallel.enable()
    import sys as _sys
    if _sys.modules["tor10_b200.UniTensor"]._dense_shard(chi * d, chi * d, chi * d, x) is None:
        same = False
    two = T.Contract(t1, t2).Storage
    parallel.disable()
    err = max(err, float((two.reshape(want.shape) - want).abs().max() / want.abs().max()))
    same = same and float((two - one).abs().max()) <= 1e-13 * float(one.abs().max())
    torch.cuda.synchronize()
    q.put((rank, "" if same else "transports disagree, or sharded result differs from the single-GPU result", err))
    dist.destroy_process_group()


def _worker_streamk(rank, world, port, q):
    """Stream-K pieces + fused all-gather in one kernel (t10_ggemm_streamk_allgather), forced on a group small
    enough for a test: tiles are cut across CTAs, parked partials are fixed up, the finished tiles are pushed to
    both ranks from the staging tile that aliases the pipeline stages."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TOR10_SHARD_MIN_MFLOP="0", TOR10_GGEMM_CFG="46",
                      TOR10_GGEMM_STREAMK="force")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import tor10_b200 as T
    from tor10_b200 import parallel
    from helpers import psym
    from test_gpu_parity import _c3_spec
    A, B = _c3_spec(1000, -7, 6, 2.5)
    a, b = psym(T, A, seed=51, device="cuda:%d" % rank), psym(T, B, seed=52, device="cuda:%d" % rank)
    single = T.Contract(a, b)
    # the k-cuts of a tile depend on the rank's share of the sectors, which differs between the transports (LPT
    # owners for the fused path, contiguous ranges for NCCL): bit-identical within a transport, rounding-level
    # agreement (1e-13) with the single-GPU stream-K result across them
    why, err, parked = [], 0.0, 0
    for transport in ("fused", "nccl", "fused"):
        parallel.enable(transport=transport)
        T._plan.clear_caches()
        first = None
        for _ in range(3):
            sharded = T.Contract(a, b)
            if first is None:
                first = [x.clone() for x in sharded.Storage]
            if not all(torch.equal(x, y) for x, y in zip(first, sharded.Storage)):
                why.append("%s: not bit-reproducible" % transport)
            err = max(err, max(float((x - y).abs().max() / x.abs().max()) for x, y in zip(single.Storage, sharded.Storage)))
        pl = T._plan.get_contract_plan(a, b)
        if not (pl.gemm.streamk and pl.gemm.tile_cfg == 46):
            why.append("%s: stream-K not in use (cfg %d)" % (transport, pl.gemm.tile_cfg))
        parked = max(parked, getattr(pl.gemm, "sk_parked", 0))
        if transport == "fused" and parallel.fused_status(pl._stage):
            why.append("fused: a wait on a peer's flag timed out")
        parallel.disable()
        T._plan.clear_caches()
    if parked == 0:
        why.append("no tile was cut across CTAs")
    torch.cuda.synchronize()
    q.put((rank, "; ".join(why), err))
    dist.destroy_process_group()


def _worker_resident(rank, world, port, q):
    """Sector-resident chaining: Contract -> Contract -> Contract with results that are never gathered.
    c = a.b stays resident; c.e reads c IN PLACE (ownership inherited from c's row sectors); relabelled c needs a
    RELAYOUT whose foreign pieces are read from the owner's arena over NVLink; x.c reads resident c as the SECOND
    operand (foreign sectors pulled).  Ten rounds wrap the ring of result arenas (write-after-read protocol).
    Everything is compared with the same chain on one GPU (rounding level) and, first link, with the oracle."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TOR10_SHARD_MIN_MFLOP="0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import tor10_b200 as T
    from tor10_b200 import parallel
    from helpers import osym, psym, rel_err
    from oracle import tor10_oracle as orc
    from test_gpu_parity import _c3_spec
    dev = "cuda:%d" % rank
    chi = 160
    A, B = _c3_spec(chi, -7, 6, 2.5)
    a, b = psym(T, A, seed=41, device=dev), psym(T, B, seed=42, device=dev)
    E = dict(B, labels=[4, 5, 6, 7])
    e = psym(T, E, seed=43, device=dev)
    vb, pb = B["bonds"][1], B["bonds"][0]                      # (chi KET), (p KET)
    E3 = {"bonds": [vb, pb, B["bonds"][2], B["bonds"][3]], "rowrank": 2, "labels": [4, 5, 6, 7]}
    e3 = psym(T, E3, seed=44, device=dev)
    A2 = {"bonds": [A["bonds"][0], A["bonds"][2], A["bonds"][3], A["bonds"][1]], "rowrank": 2, "labels": [8, 9, 0, 1]}
    a2 = psym(T, A2, seed=45, device=dev)
    why, err = [], 0.0

    def chain():
        c = T.Contract(a, b)                                   # labels [0, 1, 4, 5]
        d = T.Contract(c, e)                                   # c in place, ownership inherited
        c3 = T.Contract(a, b)
        c3.SetLabels([0, 1, 5, 4])
        f = T.Contract(c3, e3)                                 # relayout of a resident operand
        g = T.Contract(a2, c)                                  # resident SECOND operand: foreign sectors pulled
        return c, d, f, g

    ref = [[x.clone() for x in t.Storage] for t in chain()]
    oc = orc.contract_sym(osym(A, seed=41), osym(B, seed=42))

    def svd_chain():
        """Contract -> Svd_truncate -> Contract -> Contract: the sector SVDs are dealt over the ranks, the factors stay
        resident, the product U.S.V is rebuilt from them (rank-`keep` approximation of c)."""
        c = T.Contract(a, b)
        U, S, V = T.Svd_truncate(c, 120)
        R = T.Contract(T.Contract(U, S), V)
        return S, R

    svd_ref = [[x.clone() for x in t.Storage] for t in svd_chain()]

    # config-5 shape: a Network of U1 tensors (E1 - P - E2), launched twice; the intermediate stays resident
    u1 = [["U1", 0]]
    dq = sorted([[1], [0], [-1]], reverse=True)
    cq = sorted(orc.gauss_u1_qnums(12, -1, 1, 0.8).tolist(), reverse=True)

    def bd(dim, bt, qn):
        return {"dim": dim, "btype": bt, "qnums": qn, "syms": u1}
    nspecs = {"E1": {"bonds": [bd(12, -1, cq), bd(3, 1, dq), bd(12, 1, cq)], "rowrank": 1, "labels": [0, 1, 2]},
              "P": {"bonds": [bd(3, -1, dq), bd(3, -1, dq), bd(3, 1, dq), bd(3, 1, dq)], "rowrank": 2, "labels": [0, 1, 2, 3]},
              "E2": {"bonds": [bd(12, -1, cq), bd(3, -1, dq), bd(12, 1, cq)], "rowrank": 2, "labels": [0, 1, 2]}}
    net = T.Network()
    net.Fromstring("E1: 10; 11 12\nP: 11 20; 21 13\nE2: 12 13; 14\nTOUT: 10 20; 21 14\nOrder: (E1,P),E2\n")
    for i, (k, sp) in enumerate(nspecs.items()):
        net.Put(k, psym(T, sp, seed=600 + i, device=dev))

    def launch():
        net.Launch()
        return net.Launch().Contiguous()

    net_ref = [x.clone() for x in launch().Storage]
    for tr in ("resident", "pull", "push"):
        parallel.enable(transport=tr)
        for it in range(10):
            c, d, f, g = chain()
            if it == 0:
                if c._res is None or d._res is None or f._res is None or g._res is None:
                    why.append("%s: results are not resident" % tr)
                elif tr == "resident":
                    if c._res.valid.all() and world > 1:
                        why.append("resident: every sector is local (nothing was sharded)")
                    pl = T._plan.get_contract_plan(c, e)
                    if not getattr(pl, "inherited", False):
                        why.append("resident: ownership was not inherited from the resident operand")
            if it == 0:
                S, R = svd_chain()
                if tr == "resident" and (S._res is None or R._res is None):
                    why.append("resident: the factors of the sharded Svd_truncate are not resident")
                for gt, rf in zip(([x.clone() for x in S.Storage], [x.clone() for x in R.Storage]), svd_ref):
                    if len(gt) != len(rf):
                        why.append("%s: Svd_truncate kept different sectors" % tr)
                        continue
                    for x, y in zip(gt, rf):
                        den = float(y.abs().max())
                        err = max(err, float((x - y).abs().max()) / (den if den > 0 else 1.0))
            if it == 0:
                for x, y in zip(launch().Storage, net_ref):
                    den = float(y.abs().max())
                    err = max(err, float((x - y).abs().max()) / (den if den > 0 else 1.0))
            if it in (0, 9):
                got = [[x.clone() for x in t.Storage] for t in (c, d, f, g)]   # gathers on demand (every rank)
                for gt, rf in zip(got, ref):
                    for x, y in zip(gt, rf):
                        den = float(y.abs().max())
                        err = max(err, float((x - y).abs().max()) / (den if den > 0 else 1.0))
                if it == 0:
                    err = max(err, max(rel_err(x.cpu().numpy(), y) for x, y in zip(got[0], oc.blocks)))
        parallel.disable()
        T._plan.clear_caches()
    torch.cuda.synchronize()
    q.put((rank, "; ".join(why), err))
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker, _worker_streamk, _worker_resident],
                         ids=["transports", "streamk_allgather", "resident_chain"])
def test_sharded_contract_two_gpus(worker):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, why, err in res:
        assert not why, "rank %d: %s" % (rank, why)
        assert err <= 1e-12      # against the oracle (transports) / the single-GPU stream-K result (streamk_allgather)
