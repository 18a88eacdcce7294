"""Comm roofline of the sector gather: achieved NVLink bandwidth of plain peer copies between the
ranks of one box, measured with CUDA events (max over ranks) -- LSU kernel pulls and pushes (t10_gather_peers: 16-byte
vectors, all SMs; the all-gather pattern: every rank moves 1/world of the buffer from / to every peer at once) and
copy-engine transfers (Tensor.copy_ on peer mappings).
    python -m torch.distributed.run --nproc-per-node N tools/nvlink_probe.py > profiles/r02_nvlink_peak_nN.json"""
import ctypes as C, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
from tor10_b200 import _lib
lib = _lib.load()
out = {"world": world, "cases": []}
for mb in (3.125, 12.5, 25.0, 100.0):
    n = int(mb * 1e6 / 8) // (64 * world) * (64 * world)
    buf = symm_mem.empty(n, dtype=torch.float64, device=dev)
    buf.fill_(rank + 1.0)
    hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD)
    peers = [hdl.get_buffer(r, (n,), torch.float64) for r in range(world)]
    local = torch.empty(n, dtype=torch.float64, device=dev)
    sl = n // world
    order = [(rank + 1 + i) % world for i in range(world - 1)]

    def segs(ptr_of, who):
        k = len(who)
        return (k, (C.c_void_p * k)(*[ptr_of(r) for r in who]), (C.c_int64 * k)(*[r * sl for r in who]),
                (C.c_int64 * k)(*[(r + 1) * sl for r in who]))

    def pull():     # this rank reads slice r from rank r, for every peer r
        k, p, lo, hi = segs(lambda r: peers[r].data_ptr(), order)
        _lib.check(lib.t10_gather_peers(0, local.data_ptr(), k, p, lo, hi, _lib.stream_ptr()))

    def push():     # this rank writes ITS slice into every peer (one launch per peer: dst differs)
        for r in order:
            _lib.check(lib.t10_gather_peers(0, peers[r].data_ptr(), 1, (C.c_void_p * 1)(buf.data_ptr()),
                                            (C.c_int64 * 1)(rank * sl), (C.c_int64 * 1)((rank + 1) * sl), _lib.stream_ptr()))

    def ce_pull():
        for r in order:
            local[r * sl:(r + 1) * sl].copy_(peers[r][r * sl:(r + 1) * sl], non_blocking=True)

    def pair_pull():   # one peer only, the whole buffer: point-to-point rate
        r = (rank + 1) % world
        _lib.check(lib.t10_gather_peers(0, local.data_ptr(), 1, (C.c_void_p * 1)(peers[r].data_ptr()),
                                        (C.c_int64 * 1)(0), (C.c_int64 * 1)(n), _lib.stream_ptr()))

    for name, fn, nbytes in (("allgather_pull_lsu", pull, 8 * sl * (world - 1)), ("allgather_push_lsu", push, 8 * sl * (world - 1)),
                             ("allgather_pull_copy_engine", ce_pull, 8 * sl * (world - 1)), ("pair_pull_lsu", pair_pull, 8 * n)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        ts = []
        for _ in range(10):
            dist.barrier(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = torch.tensor([float(np.median(ts))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["cases"].append({"buffer_MB": 8 * n / 1e6, "op": name, "bytes_in_per_gpu": nbytes, "us": float(t.item()) * 1e3,
                             "GBps_per_gpu_per_direction": nbytes / (float(t.item()) * 1e-3) / 1e9})
    del peers, hdl, buf
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
