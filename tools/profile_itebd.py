import cProfile, pstats, importlib.util, os, sys, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
chi = 32
rng = np.random.Generator(np.random.PCG64(7))
init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
ex.run(1.0, 0.5, chi, 0.0, max_steps=20, init=init, check_every=10**9)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
ex.run(1.0, 0.5, chi, 0.0, max_steps=200, init=init, check_every=200)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
