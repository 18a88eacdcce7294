#!/usr/bin/env python
"""
Generates tests/golden/*.npz by running the UNMODIFIED Python reference
(/root/reference, kaihsin/Tor10) on seeded inputs.  Runs only in the build
container (the reference does not travel to the GPU box); the produced fixtures
are committed and are what the tests read.

    python tests/golden/make_golden.py

Compatibility shim (reference files untouched):
  np.int was removed in numpy>=1.24 (Bond.py:253 ...), torch.symeig in torch 2.x.

Every case is described by a small JSON "spec" stored next to its arrays so the
tests can rebuild the same input for the oracle and for the CUDA path:
  bond  = {"dim", "btype" (+1 BRA / -1 KET / 0 REG), "qnums" (already sorted as the
           reference sorts them, Bond.py:275-276), "syms" [["U1",0] | ["Zn",n]]}
  block values: numpy PCG64(seed).random(shape) per block in sector order.
"""
import json
import os
import sys
import warnings

warnings.filterwarnings("ignore")
import numpy as np

np.int = int
import torch

torch.symeig = lambda a, eigenvectors=False, upper=True: torch.linalg.eigh(a, UPLO="U" if upper else "L")
sys.path.insert(0, "/root/reference")
import tor10  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BT = {1: tor10.BD_BRA, -1: tor10.BD_KET, 0: tor10.BD_REG}
BTI = {tor10.BD_BRA: 1, tor10.BD_KET: -1, tor10.BD_REG: 0}


def mk_sym(s):
    return tor10.Symmetry.U1() if s[0] == "U1" else tor10.Symmetry.Zn(s[1])


def mk_bond(b):
    if b.get("qnums") is None:
        return tor10.Bond(b["dim"], BT[b["btype"]])
    return tor10.Bond(b["dim"], BT[b["btype"]], qnums=b["qnums"], sym_types=[mk_sym(s) for s in b["syms"]])


def bond_spec(bd, syms):
    return {"dim": int(bd.dim), "btype": BTI[bd.bondType], "qnums": bd.qnums.tolist(), "syms": syms}


def rand_bond(rng, syms, btype, dim=None):
    dim = int(rng.integers(1, 6)) if dim is None else dim
    cols = []
    for s in syms:
        if s[0] == "U1":
            cols.append(rng.integers(-2, 3, size=dim))
        else:
            cols.append(rng.integers(0, s[1], size=dim))
    q = np.stack(cols, axis=1)
    bd = tor10.Bond(dim, BT[btype], qnums=q.tolist(), sym_types=[mk_sym(s) for s in syms])
    return bond_spec(bd, syms)


def mk_tensor(spec):
    return tor10.UniTensor(bonds=[mk_bond(b) for b in spec["bonds"]], rowrank=spec["rowrank"],
                           labels=spec["labels"])


def fill(T, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    for b in range(len(T.Storage)):
        T.Storage[b] = torch.from_numpy(rng.random(tuple(T.Storage[b].shape)))
    return T


def dump_maps(out, key, T):
    out[key + "/block_qnums"] = np.asarray(T._block_qnums, dtype=np.int64)
    out[key + "/ket_map"] = np.asarray(T._Ket_mapper_blks, dtype=np.int64)
    out[key + "/bra_map"] = np.asarray(T._Bra_mapper_blks, dtype=np.int64)
    out[key + "/accu_in"] = np.asarray(T._accu_off_in, dtype=np.int64)
    out[key + "/accu_out"] = np.asarray(T._accu_off_out, dtype=np.int64)
    out[key + "/shapes"] = np.array([tuple(s.shape) for s in T.Storage], dtype=np.int64).reshape(-1, 2)
    for b in range(len(T.Storage)):
        out[key + "/ket_inv%d" % b] = np.asarray(T._Ket_invmapper_blks[b], dtype=np.int64)
        out[key + "/bra_inv%d" % b] = np.asarray(T._Bra_invmapper_blks[b], dtype=np.int64)


def dump_blocks(out, key, T):
    out[key + "/nblk"] = np.array(len(T.Storage))
    for b in range(len(T.Storage)):
        out[key + "/blk%d" % b] = T.Storage[b].numpy().copy()


SYMSETS = [[["U1", 0]], [["Zn", 2]], [["Zn", 3]], [["U1", 0], ["Zn", 2]], [["U1", 0], ["U1", 0]],
           [["Zn", 4], ["U1", 0], ["Zn", 2]]]


def rand_tensor_spec(rng, rank=None, syms=None):
    rank = int(rng.integers(2, 6)) if rank is None else rank
    syms = SYMSETS[int(rng.integers(0, len(SYMSETS)))] if syms is None else syms
    while True:
        bt = rng.choice([-1, 1], size=rank)
        if (bt == 1).any() and (bt == -1).any():
            break
    bonds = [rand_bond(rng, syms, int(bt[i])) for i in range(rank)]
    rowrank = int(rng.integers(1, rank))
    labels = rng.permutation(np.arange(-3, 12))[:rank].tolist()
    return {"bonds": bonds, "rowrank": rowrank, "labels": [int(x) for x in labels]}


# ---------------------------------------------------------------------------------
def gen_blockmap():
    """a8 block maps (+ 'raises' cases with no common sector)."""
    rng = np.random.default_rng(20261019)
    out, specs = {}, {}
    n_ok = n_bad = 0
    i = 0
    while n_ok < 60 or n_bad < 6:
        spec = rand_tensor_spec(rng)
        key = "bm%03d" % i
        try:
            T = mk_tensor(spec)
        except TypeError:
            if n_bad < 6:
                spec["raises"] = True
                specs[key] = spec
                n_bad += 1
                i += 1
            continue
        if n_ok >= 60:
            continue
        spec["raises"] = False
        specs[key] = spec
        dump_maps(out, key, T)
        tin, tout = T.GetTotalQnums(physical=False)
        out[key + "/tq_in"] = tin.qnums.astype(np.int64)
        out[key + "/tq_out"] = tout.qnums.astype(np.int64)
        n_ok += 1
        i += 1
    out["specs"] = np.array(json.dumps(specs))
    np.savez_compressed(os.path.join(HERE, "blockmap.npz"), **out)
    print("blockmap:", n_ok, "ok", n_bad, "raising")


def gen_relayout():
    """a9+a10: Permute -> Contiguous_ on filled tensors; and GetBlock/PutBlock through a
    non-contiguous view (UniTensor.py:3036-3053, :3255-3278)."""
    rng = np.random.default_rng(777)
    out, specs = {}, {}
    n = 0
    while n < 40:
        spec = rand_tensor_spec(rng, rank=int(rng.integers(2, 6)))
        try:
            T = mk_tensor(spec)
        except TypeError:
            continue
        rank = len(spec["bonds"])
        perm = rng.permutation(rank)
        new_rr = int(rng.integers(1, rank))
        if (perm == np.arange(rank)).all() and new_rr != spec["rowrank"]:
            continue        # quirk Q1: layout goes stale, result undefined
        seed = 5000 + n
        fill(T, seed)
        try:
            T.Permute(perm.tolist(), rowrank=new_rr)
            # after the permute the *new* split must still own a common sector and the tensor
            # must stay constructible (>=1 bra/ket on each side is not required by Permute itself)
            P = T.Contiguous()
        except (TypeError, Exception) as e:  # noqa
            continue
        key = "rl%03d" % n
        spec.update({"perm": perm.tolist(), "new_rowrank": new_rr, "seed": seed})
        out[key + "/logical_block_qnums"] = np.asarray(T._block_qnums, dtype=np.int64)
        out[key + "/mapper"] = np.asarray(T._mapper, dtype=np.int64)
        out[key + "/inv_mapper"] = np.asarray(T._inv_mapper, dtype=np.int64)
        out[key + "/contig_flag"] = np.array(bool(T._contiguous))
        dump_maps(out, key + "/out", P)
        dump_blocks(out, key + "/out", P)
        # GetBlock on the non-contiguous tensor for every logical sector
        for s, q in enumerate(np.asarray(T._block_qnums)):
            out[key + "/getblk%d" % s] = T.GetBlock(*q.tolist()).Storage.numpy().copy()
        # PutBlock through the non-contiguous view, then read the memory-layout storage back
        if len(T._block_qnums):
            q0 = np.asarray(T._block_qnums)[0]
            blk = T.GetBlock(*q0.tolist())
            vals = np.arange(blk.Storage.numel(), dtype=np.float64).reshape(tuple(blk.Storage.shape)) + 100.0
            blk.Storage = torch.from_numpy(vals)
            T.PutBlock(blk, *q0.tolist())
            dump_blocks(out, key + "/afterput", T)
        specs[key] = spec
        n += 1
    out["specs"] = np.array(json.dumps(specs))
    np.savez_compressed(os.path.join(HERE, "relayout.npz"), **out)
    print("relayout:", n)


def wd_pair(rng, syms, nfa, nk, nfb, maxdim=4):
    """A pair in the well-defined parity domain: A free = KET, contracted = BRA on A /
    KET on B (same charges), B free = BRA; storage order shuffled so a relayout is needed."""
    fa = [rand_bond(rng, syms, -1, dim=int(rng.integers(1, maxdim + 1))) for _ in range(nfa)]
    kk = [rand_bond(rng, syms, 1, dim=int(rng.integers(1, maxdim + 1))) for _ in range(nk)]
    fb = [rand_bond(rng, syms, 1, dim=int(rng.integers(1, maxdim + 1))) for _ in range(nfb)]
    kb = [dict(b, btype=-1) for b in kk]
    la = list(range(0, nfa))
    lk = list(range(20, 20 + nk))
    lb = list(range(40, 40 + nfb))
    lk = [int(x) for x in rng.permutation(lk)]      # label order != position order (quirk Q5)
    A_b, A_l = fa + kk, la + lk
    B_b, B_l = kb + fb, lk + lb
    pa = rng.permutation(len(A_b))
    pb = rng.permutation(len(B_b))
    A = {"bonds": [A_b[i] for i in pa], "labels": [A_l[i] for i in pa]}
    B = {"bonds": [B_b[i] for i in pb], "labels": [B_l[i] for i in pb]}
    for S in (A, B):
        S["rowrank"] = int(rng.integers(1, len(S["bonds"])))
    return A, B


def need_ok(spec, free_first, nfree):
    """Quirk Q1 filter: operand must need either nothing or a non-identity bond permutation."""
    labels = spec["labels"]
    common_sorted = sorted(l for l in labels if 20 <= l < 40)
    free_pos = [i for i, l in enumerate(labels) if not (20 <= l < 40)]
    com_pos = [labels.index(l) for l in common_sorted]
    perm = free_pos + com_pos if free_first else com_pos + free_pos
    rr = nfree if free_first else len(com_pos)
    ident = perm == list(range(len(labels)))
    return (not ident) or spec["rowrank"] == rr


def gen_contract_sym():
    rng = np.random.default_rng(4242)
    out, specs = {}, {}
    n = 0
    tries = 0
    while n < 40:
        tries += 1
        syms = SYMSETS[int(rng.integers(0, len(SYMSETS)))]
        nfa, nk, nfb = int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 3))
        A, B = wd_pair(rng, syms, nfa, nk, nfb)
        if not (need_ok(A, True, nfa) and need_ok(B, False, nfb)):
            continue
        try:
            TA, TB = mk_tensor(A), mk_tensor(B)
            fill(TA, 9000 + 2 * n)
            fill(TB, 9001 + 2 * n)
            C = tor10.Contract(TA, TB)
        except Exception:
            continue
        key = "cs%03d" % n
        specs[key] = {"A": A, "B": B, "seedA": 9000 + 2 * n, "seedB": 9001 + 2 * n}
        out[key + "/labels"] = np.asarray(C.labels, dtype=np.int64)
        out[key + "/braket"] = np.asarray(C.braket, dtype=np.int64)
        out[key + "/rowrank"] = np.array(int(C.rowrank))
        dump_maps(out, key + "/C", C)
        dump_blocks(out, key + "/C", C)
        # independent cross-check: dense einsum of the densified operands
        n += 1
    # docstring example UniTensor.py:3603-3676 (outside the well-defined domain, still a U1 golden)
    u1 = [["U1", 0]]
    bd1 = {"dim": 3, "btype": -1, "qnums": [[0], [1], [2]], "syms": u1}
    bd2 = {"dim": 4, "btype": -1, "qnums": [[-1], [2], [0], [2]], "syms": u1}
    bd3 = {"dim": 5, "btype": 1, "qnums": [[4], [2], [-1], [5], [1]], "syms": u1}
    for b in (bd1, bd2, bd3):
        b["qnums"] = mk_bond(b).qnums.tolist()
    bd1b = dict(bd1, btype=1)
    bd2b = dict(bd2, btype=1)
    bd4 = {"dim": 6, "btype": -1, "qnums": [[-1], [0], [2], [1], [0], [0]], "syms": u1}   # placeholder, replaced below
    A = {"bonds": [bd1, bd2, bd3], "labels": [10, 11, 12], "rowrank": 2}
    # sym_B of the docstring: labels 11,10,7 ; bonds bra(bd2), bra(bd1), and a ket bond with dim 5
    bd7 = {"dim": 5, "btype": -1, "qnums": [[1], [0], [-1], [-2], [-2]], "syms": u1}
    bd7["qnums"] = mk_bond(bd7).qnums.tolist()
    for rr in (1, 2):
        Bs = {"bonds": [bd2b, bd1b, bd7], "labels": [11, 10, 7], "rowrank": rr}
        try:
            TA, TB = fill(mk_tensor(A), 31), fill(mk_tensor(Bs), 32)
        except TypeError:
            continue
        for nm, (x, y) in (("ab", (TA, TB)), ("ba", (TB, TA))):
            try:
                C = tor10.Contract(x, y)
            except Exception:
                continue
            key = "doc_%s_rr%d" % (nm, rr)
            specs[key] = {"A": A if nm == "ab" else Bs, "B": Bs if nm == "ab" else A,
                          "seedA": 31 if nm == "ab" else 32, "seedB": 32 if nm == "ab" else 31}
            out[key + "/labels"] = np.asarray(C.labels, dtype=np.int64)
            out[key + "/braket"] = np.asarray(C.braket, dtype=np.int64)
            out[key + "/rowrank"] = np.array(int(C.rowrank))
            dump_maps(out, key + "/C", C)
            dump_blocks(out, key + "/C", C)
    out["specs"] = np.array(json.dumps(specs))
    np.savez_compressed(os.path.join(HERE, "contract_sym.npz"), **out)
    print("contract_sym:", n, "random (+doc)", "tries", tries)


def gen_dense():
    """a13 dense Contract (untagged, tagged, diag operands, outer product), a14 Matmul /
    Chain_matmul, a15 Network (test.net example and an ordered 4-tensor network)."""
    rng = np.random.default_rng(99)
    out, specs = {}, {}

    def dense_ut(dims, labels, rowrank, tagged, is_diag, seed):
        if tagged:
            bt = [tor10.BD_KET] * rowrank + [tor10.BD_BRA] * (len(dims) - rowrank)
            # shuffle tags a bit so both mismatching spaces occur; contraction needs bra<->ket
        else:
            bt = [tor10.BD_REG] * len(dims)
        bonds = [tor10.Bond(int(d), t) for d, t in zip(dims, bt)]
        T = tor10.UniTensor(bonds=bonds, rowrank=rowrank, labels=labels, is_diag=is_diag)
        r = np.random.Generator(np.random.PCG64(seed))
        T.Storage = torch.from_numpy(r.random(tuple(T.Storage.shape)))
        return T

    cases = [
        # (dimsA, labelsA, rrA, diagA, dimsB, labelsB, rrB, diagB, tagged)
        ([5, 2, 4, 3], [6, 1, 7, 8], 2, False, [4, 2, 3, 6], [7, 2, 10, 9], 2, False, False),   # docstring :3530-3598
        ([4, 2, 3, 6], [7, 2, 10, 9], 2, False, [5, 2, 4, 3], [6, 1, 7, 8], 2, False, False),
        ([6, 3, 6], [-1, 0, -2], 1, False, [6, 6], [-2, -3], 1, True, False),                    # iTEBD.py:81
        ([6, 6], [-4, -1], 1, True, [6, 3, 6], [-1, 0, -2], 1, False, False),                    # iTEBD.py:83
        ([3, 2, 2, 3], [-4, 0, 1, -5], 1, False, [2, 2, 2, 2], [0, 1, 2, 3], 2, False, False),   # iTEBD.py:94
        ([3, 2, 2, 3], [-4, 0, 1, -5], 1, False, [3, 2, 2, 3], [-4, 0, 1, -5], 1, False, False), # full contraction :93
        ([3, 4], [0, 1], 1, False, [2, 5], [2, 3], 1, False, False),                             # outer product :3825
        ([4, 4], [0, 1], 1, True, [4, 4], [1, 2], 1, True, False),                               # diag x diag
        ([3, 4, 2], [0, 1, 2], 2, False, [2, 4, 5], [2, 1, 3], 2, False, True),                  # tagged: bra<->ket pairs
        ([2, 3, 4, 2, 3], [9, 4, 1, 5, 2], 3, False, [3, 2, 4, 2], [2, 9, 1, 7], 1, False, False),
    ]
    for i, (da, la, ra, ga, db, lb, rb, gb, tagged) in enumerate(cases):
        key = "dc%02d" % i
        if tagged:
            # a: (KET,KET | BRA) ; b: (KET,KET|BRA) with contracted a-bra(2) <-> b-ket(2), a-ket(1) <-> ... needs opposite
            A = tor10.UniTensor(bonds=[tor10.Bond(3, tor10.BD_KET), tor10.Bond(4, tor10.BD_BRA), tor10.Bond(2, tor10.BD_BRA)],
                                rowrank=ra, labels=la)
            B = tor10.UniTensor(bonds=[tor10.Bond(2, tor10.BD_KET), tor10.Bond(4, tor10.BD_KET), tor10.Bond(5, tor10.BD_BRA)],
                                rowrank=rb, labels=lb)
            r = np.random.Generator(np.random.PCG64(300 + i))
            A.Storage = torch.from_numpy(r.random(tuple(A.Storage.shape)))
            B.Storage = torch.from_numpy(r.random(tuple(B.Storage.shape)))
            specs[key] = {"tagged": True, "A": {"dims": da, "labels": la, "rowrank": ra, "braket": [-1, 1, 1]},
                          "B": {"dims": db, "labels": lb, "rowrank": rb, "braket": [-1, -1, 1]}, "seed": 300 + i}
        else:
            A = dense_ut(da, la, ra, False, ga, 300 + i)
            B = dense_ut(db, lb, rb, False, gb, 400 + i)
            specs[key] = {"tagged": False,
                          "A": {"dims": da, "labels": la, "rowrank": ra, "is_diag": ga, "seed": 300 + i},
                          "B": {"dims": db, "labels": lb, "rowrank": rb, "is_diag": gb, "seed": 400 + i}}
        C = tor10.Contract(A, B)
        out[key + "/data"] = C.Storage.contiguous().numpy().copy()
        out[key + "/labels"] = np.asarray(C.labels, dtype=np.int64)
        out[key + "/rowrank"] = np.array(int(C.rowrank))
        if C.braket is not None:
            out[key + "/braket"] = np.asarray(C.braket, dtype=np.int64)

    # Matmul / Chain_matmul
    mats = []
    for j, (m, k) in enumerate([(3, 4), (4, 5), (5, 6), (6, 2)]):
        T = dense_ut([m, k], [0, 1], 1, False, False, 600 + j)
        mats.append(T)
    out["mm/ab"] = tor10.Matmul(mats[0], mats[1]).Storage.numpy().copy()
    out["mm/chain"] = tor10.Chain_matmul(*mats).Storage.numpy().copy()
    dg = dense_ut([4, 4], [0, 1], 1, False, True, 610)
    out["mm/a_diag"] = tor10.Matmul(mats[0], dg).Storage.numpy().copy()
    out["mm/diag_b"] = tor10.Matmul(dg, mats[1]).Storage.numpy().copy()
    dd = tor10.Matmul(dg, dg)
    out["mm/diag_diag"] = dd.Storage.numpy().copy()           # scalar dot product (reference bug)
    out["mm/diag_diag_isdiag"] = np.array(bool(dd.is_diag))

    # Network: test.net example  (Network.py:216-287)
    net_txt = open("/root/reference/test.net").read()
    specs["net0"] = {"text": net_txt, "tensors": {"A": [[3, 4, 2, 5], 2, 700], "B": [[2, 6], 1, 701], "C": [[5, 3], 1, 702]}}
    N = tor10.Network()
    N.Fromfile("/root/reference/test.net")
    for nm, (dims, rr, seed) in specs["net0"]["tensors"].items():
        N.Put(nm, dense_ut(dims, list(range(len(dims))), rr, False, False, seed))
    R = N.Launch()
    out["net0/data"] = R.Storage.contiguous().numpy().copy()
    out["net0/labels"] = np.asarray(R.labels, dtype=np.int64)
    out["net0/rowrank"] = np.array(int(R.rowrank))
    # same network without the Order line (sequential fold, Network.py:404-415)
    import tempfile
    txt2 = "\n".join(l for l in net_txt.splitlines() if not l.strip().startswith("Order"))
    with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
        f.write(txt2)
    specs["net1"] = {"text": txt2, "tensors": specs["net0"]["tensors"]}
    N = tor10.Network()
    N.Fromfile(f.name)
    for nm, (dims, rr, seed) in specs["net1"]["tensors"].items():
        N.Put(nm, dense_ut(dims, list(range(len(dims))), rr, False, False, seed))
    R = N.Launch()
    out["net1/data"] = R.Storage.contiguous().numpy().copy()
    out["net1/labels"] = np.asarray(R.labels, dtype=np.int64)
    out["net1/rowrank"] = np.array(int(R.rowrank))
    os.unlink(f.name)
    # an ordered 4-tensor ring with a nested order string
    txt3 = "W: 1; 2 3\nX: 3; 4\nY: 4 5; 6\nZ: 2 6; 5 7\nTOUT: 1; 7\nOrder: (W,(X,Y)),Z\n"
    with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
        f.write(txt3)
    specs["net2"] = {"text": txt3, "tensors": {"W": [[3, 4, 2], 1, 710], "X": [[2, 5], 1, 711],
                                               "Y": [[5, 3, 6], 2, 712], "Z": [[4, 6, 3, 2], 2, 713]}}
    N = tor10.Network()
    N.Fromfile(f.name)
    for nm, (dims, rr, seed) in specs["net2"]["tensors"].items():
        N.Put(nm, dense_ut(dims, list(range(len(dims))), rr, False, False, seed))
    R = N.Launch()
    out["net2/data"] = R.Storage.contiguous().numpy().copy()
    out["net2/labels"] = np.asarray(R.labels, dtype=np.int64)
    out["net2/rowrank"] = np.array(int(R.rowrank))
    os.unlink(f.name)

    out["specs"] = np.array(json.dumps(specs))
    np.savez_compressed(os.path.join(HERE, "dense.npz"), **out)
    print("dense: %d contract cases + matmul + 3 networks" % len(cases))


def gen_bond_kats():
    """a2/a3/a5/a6 known answers straight from the reference (Bond sort, combine, unique, intersect)."""
    rng = np.random.default_rng(5)
    out, specs = {}, {}
    for i in range(12):
        syms = SYMSETS[i % len(SYMSETS)]
        raw_a = [[int(rng.integers(-3, 4)) if s[0] == "U1" else int(rng.integers(0, s[1])) for s in syms]
                 for _ in range(int(rng.integers(1, 7)))]
        raw_b = [[int(rng.integers(-3, 4)) if s[0] == "U1" else int(rng.integers(0, s[1])) for s in syms]
                 for _ in range(int(rng.integers(1, 7)))]
        a = tor10.Bond(len(raw_a), tor10.BD_KET, qnums=raw_a, sym_types=[mk_sym(s) for s in syms])
        b = tor10.Bond(len(raw_b), tor10.BD_KET, qnums=raw_b, sym_types=[mk_sym(s) for s in syms])
        key = "bd%02d" % i
        specs[key] = {"syms": syms, "raw_a": raw_a, "raw_b": raw_b}
        out[key + "/a_sorted"] = a.qnums.astype(np.int64)
        out[key + "/b_sorted"] = b.qnums.astype(np.int64)
        out[key + "/a_unique"] = a.GetUniqueQnums().astype(np.int64)
        out[key + "/comm"] = tor10.Bond._fx_GetCommRows(a.GetUniqueQnums(), b.GetUniqueQnums()).astype(np.int64) \
            if hasattr(tor10.Bond, "_fx_GetCommRows") else \
            sys.modules["tor10.Bond"]._fx_GetCommRows(a.GetUniqueQnums(), b.GetUniqueQnums()).astype(np.int64)
        import copy
        c = copy.deepcopy(a)
        c.combine(b)
        out[key + "/combined"] = c.qnums.astype(np.int64)
    out["specs"] = np.array(json.dumps(specs))
    np.savez_compressed(os.path.join(HERE, "bond.npz"), **out)
    print("bond: 12")


def gen_itebd():
    """Config 1 (iTEBD.py:75-136) driven through the unmodified reference library from a seeded numpy
    initial state: the per-step energies are the golden trajectory for oracle/itebd_numpy.py and for the
    product's examples/itebd.py."""
    import copy
    chi, J, Hx, nsteps = 8, 1.0, 0.5, 40
    rng = np.random.Generator(np.random.PCG64(1234))
    init = {"A": rng.random((chi, 2, chi)), "B": rng.random((chi, 2, chi)), "la": rng.random(chi), "lb": rng.random(chi)}
    Tt = tor10
    dev = torch.device("cpu")

    def op(e):
        t = Tt.UniTensor(bonds=[Tt.Bond(2), Tt.Bond(2)], rowrank=1, dtype=torch.float64, device=dev)
        t.SetElem(e)
        return t
    Sz, Sx, I = op([1, 0, 0, -1]) * J, op([0, 1, 1, 0]) * Hx, op([1, 0, 0, 1])
    H = Tt.Otimes(Sx, I) + Tt.Otimes(I, Sx) + Tt.Otimes(Sz, Sz)
    H = H.Reshape([4, 4], new_labels=[0, 1], rowrank=1)
    eH = Tt.ExpH(H * -0.1).Reshape([2, 2, 2, 2], new_labels=[0, 1, 2, 3], rowrank=2)
    H = H.Reshape([2, 2, 2, 2], new_labels=[0, 1, 2, 3], rowrank=2)
    A = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(2), Tt.Bond(chi)], rowrank=1, labels=[-1, 0, -2])
    B = Tt.UniTensor(bonds=A.bonds, rowrank=1, labels=[-3, 1, -4])
    la = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(chi)], rowrank=1, labels=[-2, -3], is_diag=True)
    lb = Tt.UniTensor(bonds=[Tt.Bond(chi), Tt.Bond(chi)], rowrank=1, labels=[-4, -5], is_diag=True)
    for t, k in ((A, "A"), (B, "B"), (la, "la"), (lb, "lb")):
        t.Storage = torch.from_numpy(init[k].copy())
    energies = []
    for i in range(nsteps):
        A.SetLabels([-1, 0, -2]); B.SetLabels([-3, 1, -4]); la.SetLabels([-2, -3]); lb.SetLabels([-4, -5])
        X = Tt.Contract(Tt.Contract(A, la), Tt.Contract(B, lb))
        lb.SetLabel(-1, idx=1)
        X = Tt.Contract(lb, X)
        Xt = copy.deepcopy(X)
        XNorm = Tt.Contract(X, Xt)
        XH = Tt.Contract(X, H)
        XH.SetLabels([-4, -5, 0, 1])
        XHX = Tt.Contract(Xt, XH)
        XeH = Tt.Contract(X, eH)
        energies.append((XHX.Storage / XNorm.Storage).item())
        XeH.Permute([-4, 2, 3, -5], by_label=True)
        XeH.Reshape_([chi * 2, chi * 2], rowrank=1)
        A, la, B = Tt.Svd_truncate(XeH, chi)
        la *= la.Norm() ** -1
        A = A.Reshape([chi, 2, chi], new_labels=[-1, 0, -2], rowrank=1)
        B = B.Reshape([chi, 2, chi], new_labels=[-3, 1, -4], rowrank=1)
        lb_inv = Tt.Inverse(lb)
        A = Tt.Contract(lb_inv, A)
        B = Tt.Contract(B, lb_inv)
        A, B = B, A
        la, lb = lb, la
    out = dict(init)
    out["energies"] = np.array(energies)
    out["specs"] = np.array(json.dumps({"chi": chi, "J": J, "Hx": Hx, "nsteps": nsteps}))
    np.savez_compressed(os.path.join(HERE, "itebd.npz"), **out)
    print("itebd: %d steps, E[-1] = %.9f" % (nsteps, energies[-1]))


if __name__ == "__main__":
    gen_bond_kats()
    gen_blockmap()
    gen_relayout()
    gen_contract_sym()
    gen_dense()
    gen_itebd()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
