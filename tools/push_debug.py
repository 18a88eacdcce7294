import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["TOR10_SHARD_MIN_MFLOP"] = "0"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = "cuda:%d" % rank
dist.init_process_group("nccl", device_id=torch.device(dev))
import tor10_b200 as T
from tor10_b200 import parallel
from helpers import psym
from test_gpu_parity import _c3_spec
A, B = _c3_spec(160, -7, 6, 2.5)
a, b = psym(T, A, seed=41, device=dev), psym(T, B, seed=42, device=dev)
e = psym(T, dict(B, labels=[4, 5, 6, 7]), seed=43, device=dev)
E3 = {"bonds": [B["bonds"][1], B["bonds"][0], B["bonds"][2], B["bonds"][3]], "rowrank": 2, "labels": [4, 5, 6, 7]}
e3 = psym(T, E3, seed=44, device=dev)
A2 = {"bonds": [A["bonds"][0], A["bonds"][2], A["bonds"][3], A["bonds"][1]], "rowrank": 2, "labels": [8, 9, 0, 1]}
a2 = psym(T, A2, seed=45, device=dev)
from oracle import tor10_oracle as orc
u1 = [["U1", 0]]
dq = sorted([[1], [0], [-1]], reverse=True)
cq = sorted(orc.gauss_u1_qnums(12, -1, 1, 0.8).tolist(), reverse=True)
def bd(dim, bt, qn): return {"dim": dim, "btype": bt, "qnums": qn, "syms": u1}
nspecs = {"E1": {"bonds": [bd(12, -1, cq), bd(3, 1, dq), bd(12, 1, cq)], "rowrank": 1, "labels": [0, 1, 2]},
          "P": {"bonds": [bd(3, -1, dq), bd(3, -1, dq), bd(3, 1, dq), bd(3, 1, dq)], "rowrank": 2, "labels": [0, 1, 2, 3]},
          "E2": {"bonds": [bd(12, -1, cq), bd(3, -1, dq), bd(12, 1, cq)], "rowrank": 2, "labels": [0, 1, 2]}}
net = T.Network()
net.Fromstring("E1: 10; 11 12\nP: 11 20; 21 13\nE2: 12 13; 14\nTOUT: 10 20; 21 14\nOrder: (E1,P),E2\n")
for i, (k, sp) in enumerate(nspecs.items()):
    net.Put(k, psym(T, sp, seed=600 + i, device=dev))
def chain():
    c = T.Contract(a, b); d = T.Contract(c, e)
    c3 = T.Contract(a, b); c3.SetLabels([0, 1, 5, 4]); f = T.Contract(c3, e3)
    g = T.Contract(a2, c)
    c4 = T.Contract(a, b)
    U, S, V = T.Svd_truncate(c4, 120)
    US = T.Contract(U, S)
    R = T.Contract(US, V)
    net.Launch()
    N = net.Launch().Contiguous()
    return {"c": c, "d": d, "f": f, "g": g, "S": S, "US": US, "R": R, "N": N}
ref = {k: [x.clone() for x in t.Storage] for k, t in chain().items()}
parallel.enable(transport=os.environ.get("TR", "push"))
out = {}
for it in range(3):
    res = chain()
    info = {k: (t._res is not None, bool(t._res.valid.all()) if t._res is not None else None, type(T._plan).__name__) for k, t in res.items()}
    for k, t in res.items():
        err = 0.0
        for x, y in zip(t.Storage, ref[k]):
            den = float(y.abs().max()); err = max(err, float((x - y).abs().max()) / (den if den > 0 else 1.0))
        out["it%d_%s" % (it, k)] = err
pl = T._plan.get_contract_plan(a, b)
out["tma"] = bool(pl.gemm.tma); out["pool"] = pl._pool is not None and pl._pool is not False
if rank == 0: print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
