#!/usr/bin/env python
"""iTEBD ground state of the transverse-field Ising chain  H = J SzSz - Hx Sx  with the tor10 API
(BASELINE config 1: untagged UniTensors, chi=32, float64).  Same algorithm and the same library calls as
the reference's iTEBD.py (two-site imaginary-time gate, Contract x10, Svd_truncate, Inverse per step),
written against `tor10_b200`; `run()` is importable so bench/tests can time it and inspect the result.

    python examples/itebd.py <J> <Hx> <chi> <converge> [max_steps]
"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tor10_b200 as Tt  # noqa: E402


def two_site_gate(J, Hx, dt=0.1, device=torch.device("cpu")):
    def op(elems):
        t = Tt.UniTensor(bonds=[Tt.Bond(2), Tt.Bond(2)], rowrank=1, dtype=torch.float64, device=device)
        t.SetElem(elems)
        return t
    Sz, Sx, Id = op([1, 0, 0, -1]) * J, op([0, 1, 1, 0]) * Hx, op([1, 0, 0, 1])
    H = Tt.Otimes(Sx, Id) + Tt.Otimes(Id, Sx) + Tt.Otimes(op([1, 0, 0, -1]) * J, op([1, 0, 0, -1]))
    del Sz
    H = H.Reshape([4, 4], new_labels=[0, 1], rowrank=1)
    eH = Tt.ExpH(H * -dt).Reshape([2, 2, 2, 2], new_labels=[0, 1, 2, 3], rowrank=2)
    return H.Reshape([2, 2, 2, 2], new_labels=[0, 1, 2, 3], rowrank=2), eH


def initial_state(chi, device=torch.device("cpu"), init=None):
    """Random (or given: dict of numpy arrays A,B,la,lb) infinite MPS  --A-la-B-lb--."""
    A = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(2), Tt.Bond(chi)], rowrank=1, labels=[-1, 0, -2], device=device).Rand()
    B = Tt.UniTensor(bonds=A.bonds, rowrank=1, labels=[-3, 1, -4], device=device).Rand()
    la = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(chi)], rowrank=1, labels=[-2, -3], is_diag=True, device=device).Rand()
    lb = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(chi)], rowrank=1, labels=[-4, -5], is_diag=True, device=device).Rand()
    if init is not None:
        for t, k in ((A, "A"), (B, "B"), (la, "la"), (lb, "lb")):
            t.Storage = torch.from_numpy(np.ascontiguousarray(init[k])).to(t.Storage.device)
    return A, B, la, lb


def step(A, B, la, lb, H, eH, chi):
    """One iTEBD update; returns the new (A, B, la, lb) (already swapped for the next bond) and the energy."""
    A.SetLabels([-1, 0, -2])
    B.SetLabels([-3, 1, -4])
    la.SetLabels([-2, -3])
    lb.SetLabels([-4, -5])
    X = Tt.Contract(Tt.Contract(A, la), Tt.Contract(B, lb))
    lb.SetLabel(-1, idx=1)
    X = Tt.Contract(lb, X)                       # (-4) --lb-A-la-B-lb-- (-5), physical legs 0, 1
    Xt = copy.deepcopy(X)
    XNorm = Tt.Contract(X, Xt)
    XH = Tt.Contract(X, H)
    XH.SetLabels([-4, -5, 0, 1])
    XHX = Tt.Contract(Xt, XH)
    XeH = Tt.Contract(X, eH)
    E = XHX.Storage / XNorm.Storage              # 0-d device tensor; the caller decides when to sync
    XeH.Permute([-4, 2, 3, -5], by_label=True)
    XeH.Reshape_([chi * 2, chi * 2], rowrank=1)
    A, la, B = Tt.Svd_truncate(XeH, chi)
    la *= la.Norm() ** -1
    A = A.Reshape([chi, 2, chi], new_labels=[-1, 0, -2], rowrank=1)
    B = B.Reshape([chi, 2, chi], new_labels=[-3, 1, -4], rowrank=1)
    lb_inv = Tt.Inverse(lb)
    A = Tt.Contract(lb_inv, A)
    B = Tt.Contract(B, lb_inv)
    return B, A, lb, la, E                       # translation symmetry: exchange the two sites


def run(J, Hx, chi, cvg, max_steps=100000, init=None, verbose=False, check_every=1):
    dev = torch.device("cpu")                    # as the reference script; tor10_b200 places it on the GPU
    H, eH = two_site_gate(J, Hx, device=dev)
    A, B, la, lb = initial_state(chi, dev, init)
    Elast, energies = 0.0, []
    for i in range(max_steps):
        A, B, la, lb, E = step(A, B, la, lb, H, eH, chi)
        if (i + 1) % check_every:
            continue
        E = E.item()
        energies.append(E)
        if verbose:
            print("Energy = {:.6f}".format(E))
        if abs(E - Elast) < cvg:
            break
        Elast = E
    return {"steps": i + 1, "E": energies[-1] if energies else float("nan"), "energies": energies,
            "state": (A, B, la, lb)}


def run_graphed(J, Hx, chi, cvg, max_steps=100000, init=None, check_every=1, verbose=False):
    """The same algorithm with the step captured into ONE CUDA graph (tor10_b200.graph.StaticLoop): ~45 launches per
    step become one graph launch.  `check_every` steps apart the energy is read back and the convergence test runs."""
    from tor10_b200.graph import StaticLoop
    dev = torch.device("cpu")
    H, eH = two_site_gate(J, Hx, device=dev)
    A, B, la, lb = initial_state(chi, dev, init)
    warm = 3
    energies_warm = []

    def counted(A, B, la, lb, H, eH, chi):
        out = step(A, B, la, lb, H, eH, chi)
        energies_warm.append(out[4])
        return out

    loop = StaticLoop(counted, state=(A, B, la, lb), consts=(H, eH, chi), warmup=warm)
    energies = [float(e.item()) for e in energies_warm[:warm]]       # the eager warm-up steps are steps of the run
    Elast = energies[-2] if len(energies) > 1 else 0.0
    i = warm - 1
    if abs(energies[-1] - Elast) >= cvg:
        Elast = energies[-1]
        for i in range(warm, max_steps):
            (E,) = loop()
            if (i + 1) % check_every:
                continue
            E = float(E.item())
            energies.append(E)
            if verbose:
                print("Energy = {:.6f}".format(E))
            if abs(E - Elast) < cvg:
                break
            Elast = E
    return {"steps": i + 1, "E": energies[-1], "energies": energies, "state": loop.state}


if __name__ == "__main__":
    if len(sys.argv) < 5:
        print(".py <J> <Hx> <chi> <converge> [max_steps]")
        sys.exit(1)
    out = run(float(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]),
              int(sys.argv[5]) if len(sys.argv) > 5 else 100000, verbose=True)
    print("[Converged!]" if out["steps"] < 100000 else "[Stopped]", "steps", out["steps"], "E = %.6f" % out["E"])
