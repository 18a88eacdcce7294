"""Shared builders: one JSON spec -> oracle object / product object with identical bits."""
import numpy as np

from oracle import tor10_oracle as orc


def obond(b):
    if b.get("qnums") is None:
        return orc.OBond(b["dim"], b["btype"])
    return orc.OBond(b["dim"], b["btype"], qnums=b["qnums"], syms=[tuple(s) for s in b["syms"]])


def osym(spec, seed=None):
    t = orc.OSym([obond(b) for b in spec["bonds"]], rowrank=spec["rowrank"], labels=spec["labels"])
    if seed is not None:
        orc.fill_blocks(t, seed)
    return t


def block_values(shapes, seed, dtype=np.float64):
    rng = np.random.Generator(np.random.PCG64(seed))
    return [rng.random(tuple(int(x) for x in s)).astype(dtype) for s in shapes]


def odense(d, tagged_braket=None):
    rng = np.random.Generator(np.random.PCG64(d["seed"]))
    shape = (d["dims"][0],) if d.get("is_diag") else tuple(d["dims"])
    return orc.ODense(rng.random(shape), d["labels"], d["rowrank"], is_diag=bool(d.get("is_diag")),
                      braket=tagged_braket)


def rel_err(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    den = np.abs(ref).max()
    return float(np.abs(x - ref).max() / (den if den > 0 else 1.0))


# --- product-side builders (tor10_b200 mirrors the reference constructor signatures) ---
def pbond(T, b):
    bt = {1: T.BD_BRA, -1: T.BD_KET, 0: T.BD_REG}[b["btype"]]
    if b.get("qnums") is None:
        return T.Bond(b["dim"], bt)
    syms = [T.Symmetry.U1() if s[0] == "U1" else T.Symmetry.Zn(s[1]) for s in b["syms"]]
    return T.Bond(b["dim"], bt, qnums=b["qnums"], sym_types=syms)


def psym(T, spec, seed=None, device="cuda", dtype=None):
    import torch
    dtype = torch.float64 if dtype is None else dtype
    t = T.UniTensor(bonds=[pbond(T, b) for b in spec["bonds"]], rowrank=spec["rowrank"], labels=spec["labels"],
                    device=torch.device(device), dtype=dtype)
    if seed is not None:
        vals = block_values([tuple(s.shape) for s in t.Storage], seed)
        for b, v in enumerate(vals):
            t.Storage[b] = torch.from_numpy(v).to(dtype)
    return t
