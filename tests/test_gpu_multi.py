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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
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
    torch.cuda.synchronize()
    q.put((rank, "" if same else "transports disagree, or sharded result differs from the single-GPU result", err))
    dist.destroy_process_group()


def _worker_streamk(rank, world, port, q):
    """Stream-K pieces + fused all-gather in one kernel (t10_ggemm_streamk_allgather), forced on a group small
    enough for a test: tiles are cut across CTAs, parked partials are fixed up, the finished tiles are pushed to
    both ranks from the staging tile that aliases the pipeline stages."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TOR10_GGEMM_CFG="46",
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


@pytest.mark.parametrize("worker", [_worker, _worker_streamk], ids=["transports", "streamk_allgather"])
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
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, why, err in res:
        assert not why, "rank %d: %s" % (rank, why)
        assert err <= 1e-12      # against the oracle (transports) / the single-GPU stream-K result (streamk_allgather)
