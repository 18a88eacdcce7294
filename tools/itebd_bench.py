"""iTEBD (BASELINE config 1: chi=32, fp64) steps per second: eager drop-in script path, CUDA-graph path, numpy port on
the host.  Prints one JSON object."""
import importlib.util, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
from oracle import itebd_numpy as it
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rng = np.random.Generator(np.random.PCG64(7))
init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
out = {"chi": chi, "steps": n}
ex.run(1.0, 0.5, chi, 0.0, max_steps=30, init=init)
torch.cuda.synchronize(); t0 = time.perf_counter()
r = ex.run(1.0, 0.5, chi, 0.0, max_steps=n, init=init)           # .item() every step, like the reference script
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["eager_steps_per_s"] = n / dt; out["eager_ms_per_step"] = dt / n * 1e3; out["E_eager"] = r["E"]
t0 = time.perf_counter()
r2 = ex.run_graphed(1.0, 0.5, chi, 0.0, max_steps=n, init=init)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["graphed_steps_per_s"] = n / dt; out["graphed_ms_per_step"] = dt / n * 1e3; out["E_graphed"] = r2["E"]
t0 = time.perf_counter()
r2 = ex.run_graphed(1.0, 0.5, chi, 0.0, max_steps=n, init=init, check_every=50)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["graphed_check50_steps_per_s"] = n / dt
t0 = time.perf_counter()
r3 = it.run(1.0, 0.5, chi, 0.0, max_steps=n, init=init)
dt = time.perf_counter() - t0
out["numpy_port_steps_per_s"] = n / dt; out["E_numpy"] = r3["E"]; out["cores"] = os.cpu_count()
out["max_abs_dE_eager_vs_numpy"] = float(np.abs(np.array(r["energies"]) - np.array(r3["energies"])).max())
print(json.dumps(out))
