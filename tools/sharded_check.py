"""Full-size check of a sector-sharded Contract under torchrun (bench workload, chi = 4096): the gathered result of
every transport equals the single-GPU result within 1e-13 and is bit-reproducible over repeated calls.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py
"""
import os, sys, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import tor10_b200 as T
from tor10_b200 import _plan, parallel
import bench
from helpers import psym

A, B = bench.c3a_specs(int(os.environ.get("CHI", "4096")))
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
single = [x.clone() for x in T.Contract(a, b).Storage]
out = {}
for tr in os.environ.get("TRANSPORTS", "resident,push,pull,fused,nccl").split(","):
    parallel.enable(transport=tr)
    _plan.clear_caches()
    first, repro, err = None, True, 0.0
    for _ in range(int(os.environ.get("REPS", "20"))):
        c = T.Contract(a, b)
        if first is None:
            first = [x.clone() for x in c.Storage]
        repro = repro and all(torch.equal(x, y) for x, y in zip(first, c.Storage))
        err = max(err, max(float((x - y).abs().max() / x.abs().max()) for x, y in zip(single, c.Storage)))
    g = _plan.get_contract_plan(a, b).gemm
    out[tr] = {"bit_reproducible": repro, "max_rel_err_vs_single": err, "tile_cfg": g.tile_cfg, "stream_k": bool(g.streamk), "tma": bool(getattr(g, "tma", False))}
    parallel.disable()
    _plan.clear_caches()
res = [None] * dist.get_world_size()
dist.all_gather_object(res, out)
if rank == 0:
    print(json.dumps({"world": dist.get_world_size(), "ranks": res}))
    ok = all(v["bit_reproducible"] and v["max_rel_err_vs_single"] <= 1e-13 for r in res for v in r.values())
    print("SHARDED_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
