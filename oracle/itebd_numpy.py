"""oracle/itebd_numpy.py -- TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT CODE.

numpy restatement of the reference's iTEBD driver (iTEBD.py:27-136: transverse-field Ising chain,
two-site gate exp(-0.1 H), 10 contractions + Svd_truncate + Inverse per step), each contraction written as
the tensordot the reference's dense Contract performs (UniTensor.py:3769) with the rows-first output order
(:3773-3784).  Pinned against the reference run in the build container: `iTEBD.py 1.0 0.5 32 1e-8`
converges to E = -1.252868; tests/test_oracle_golden.py::test_itebd_oracle_energy checks that.
"""
import numpy as np


def gate(J, Hx, dt=0.1):
    Sz = np.array([[1.0, 0.0], [0.0, -1.0]]) * J
    Sx = np.array([[0.0, 1.0], [1.0, 0.0]]) * Hx
    Id = np.eye(2)
    H = np.kron(Sx, Id) + np.kron(Id, Sx) + np.kron(Sz, Sz)           # iTEBD.py:38-46 (Otimes = kron)
    w, u = np.linalg.eigh(-dt * H)                                     # linalg.py:304-308
    eH = (u * np.exp(w)) @ u.T
    return H.reshape(2, 2, 2, 2), eH.reshape(2, 2, 2, 2)


def step(A, B, la, lb, H, eH, chi):
    """A,B: [chi,2,chi]; la,lb: [chi] diagonals.  iTEBD.py:76-133"""
    Ala = A * la[None, None, :]                                        # :81 Contract(A,la), diag densified
    Blb = B * lb[None, None, :]
    X = np.tensordot(Ala, Blb, axes=([2], [0]))                        # [-1,0,1,-5]
    X = lb[:, None, None, None] * X                                    # :83 Contract(lb,X): [-4,0,1,-5]
    norm = np.tensordot(X, X, axes=4)                                  # :93
    XH = np.tensordot(X, H, axes=([1, 2], [0, 1]))                     # :94 -> [-4,-5,2,3]
    XHX = np.tensordot(X, XH, axes=([0, 3, 1, 2], [0, 1, 2, 3]))       # :96-97 after relabel [-4,-5,0,1]
    XeH = np.tensordot(X, eH, axes=([1, 2], [0, 1]))                   # :98 -> [-4,-5,2,3]
    E = XHX / norm
    M = XeH.transpose(0, 2, 3, 1).reshape(2 * chi, 2 * chi)            # :107-112
    u, s, vh = np.linalg.svd(M, full_matrices=False)                   # :114
    u, s, vh = u[:, :chi], s[:chi], vh[:chi, :]
    s = s / np.linalg.norm(s)                                          # :116
    A2 = u.reshape(chi, 2, chi)
    B2 = vh.reshape(chi, 2, chi)
    A2 = (1.0 / lb)[:, None, None] * A2                                # :129-130
    B2 = B2 * (1.0 / lb)[None, None, :]                                # :131
    return B2, A2, lb, s, float(E)                                     # :134-135 swap


def run(J, Hx, chi, cvg, max_steps=100000, init=None, seed=0):
    H, eH = gate(J, Hx)
    if init is None:
        rng = np.random.Generator(np.random.PCG64(seed))
        init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi),
                "lb": rng.random(chi)}
    A, B, la, lb = init["A"].copy(), init["B"].copy(), init["la"].copy(), init["lb"].copy()
    Elast, energies = 0.0, []
    for i in range(max_steps):
        A, B, la, lb, E = step(A, B, la, lb, H, eH, chi)
        energies.append(E)
        if abs(E - Elast) < cvg:
            break
        Elast = E
    return {"steps": i + 1, "E": energies[-1], "energies": energies, "init": init}
