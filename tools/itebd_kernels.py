import importlib.util, os, sys
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
chi = 32
rng = np.random.Generator(np.random.PCG64(7))
init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
ex.run(1.0, 0.5, chi, 0.0, max_steps=20, init=init, check_every=10**9)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ex.run(1.0, 0.5, chi, 0.0, max_steps=20, init=init, check_every=10**9)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
