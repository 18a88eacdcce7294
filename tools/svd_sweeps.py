"""Jacobi sweeps per Svd_truncate call along an iTEBD run (warm start on / off)."""
import importlib.util, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(ROOT, "examples", "itebd.py"))
ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
from tor10_b200 import linalg
chi = 32
rng = np.random.Generator(np.random.PCG64(1234))
init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
for warm in (True, False):
    linalg._SVD_WARM = warm
    linalg._SVD_STATE.clear()
    linalg._SVD_DEBUG[:] = [torch.zeros(1, dtype=torch.int32, device="cuda"), []]
    ex.run(1.0, 0.5, chi, 0.0, max_steps=2000, init=init)
    sw = np.array(linalg._SVD_DEBUG[1])
    print(json.dumps({"warm": warm, "calls": len(sw), "mean_sweeps": float(sw.mean()), "first100": float(sw[:100].mean()),
                      "steps_500_1000": float(sw[500:1000].mean()), "last500": float(sw[-500:].mean()),
                      "hist": np.bincount(sw).tolist()}))
