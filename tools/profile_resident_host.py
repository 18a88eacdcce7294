"""cProfile of the host side of a sector-resident Contract (torchrun, 2+ ranks): where the ~50 us per call go."""
import cProfile, io, os, pstats, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
import tor10_b200 as T
from tor10_b200 import parallel
import bench
from helpers import psym
parallel.enable(transport="resident")
A, B = bench.c3a_specs(int(os.environ.get("CHI", "1024")))
a, b = psym(T, A, seed=1003, device=dev), psym(T, B, seed=1004, device=dev)
keep = [None]
for _ in range(20): keep[0] = T.Contract(a, b)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(2000): keep[0] = T.Contract(a, b)
t1 = time.perf_counter(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): keep[0] = T.Contract(a, b)
torch.cuda.synchronize(); pr.disable()
if rank == 0:
    print("host us/call %.1f" % ((t1 - t0) / 2000 * 1e6))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
dist.barrier(); dist.destroy_process_group()
