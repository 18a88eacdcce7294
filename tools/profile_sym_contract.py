import cProfile, pstats, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from helpers import psym
from bench import c3a_specs
dev = torch.device("cuda", 0)
A, B = c3a_specs(256)
a, b = psym(T, A, seed=1, device=dev), psym(T, B, seed=2, device=dev)
for _ in range(20): T.Contract(a, b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): T.Contract(a, b)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host us/call %.1f  with drain %.1f" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): T.Contract(a, b)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
