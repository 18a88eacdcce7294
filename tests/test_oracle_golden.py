"""CPU: pins the oracle (oracle/tor10_oracle.py) against
  * the known answers of the reference's own tests and docstrings,
  * fixtures produced by the unmodified reference (tests/golden/make_golden.py).
Integer maps exact; fp64 exact here too (pure data movement / same numpy matmul order is
not guaranteed, so values use 1e-13)."""
import os

import numpy as np
import pytest

from oracle import tor10_oracle as orc
from helpers import obond, odense, osym, rel_err


# ---- reference tests/test_UniTensor.py:86-111 -------------------------------------
def test_ref_GetTotalQnums_kat():
    b1 = orc.OBond(3, orc.KET, qnums=[[0, 2, 1, 0], [1, 1, -1, 1], [2, -1, 1, 0]])
    b2 = orc.OBond(4, orc.KET, qnums=[[-1, 0, -1, 3], [0, 0, -1, 2], [1, 0, 1, 0], [2, -2, -1, 1]])
    b3 = orc.OBond(2, orc.BRA, qnums=[[-4, 3, 0, -1], [1, 1, -2, 3]])
    tq_in, _ = orc.total_qnums([b1, b2, b3], [b1.btype, b2.btype, b3.btype], 2)
    want = orc.OBond(12, orc.KET, qnums=[[+4, -3, +0, +1], [+3, -1, +2, +0], [+2, -1, +0, +2],
                                         [+1, -1, +0, +3], [+3, -1, -2, +2], [+2, +1, +0, +1],
                                         [+1, +1, -2, +3], [+0, +1, -2, +4], [+2, +0, +0, +1],
                                         [+1, +2, +2, +0], [+0, +2, +0, +2], [-1, +2, +0, +3]]).qnums
    got = tq_in[np.lexsort(tq_in.T[::-1])[::-1]]
    assert got.tolist() == want.tolist()


def _t3():
    b1 = orc.OBond(3, orc.KET, qnums=[[0], [1], [2]])
    b2 = orc.OBond(4, orc.KET, qnums=[[-1], [2], [0], [2]])
    b3 = orc.OBond(5, orc.BRA, qnums=[[4], [2], [2], [5], [1]])
    return orc.OSym([b1, b2, b3], rowrank=2, labels=[10, 11, 12])


# ---- reference tests/test_UniTensor.py:113-119, :121-137 --------------------------
def test_ref_GetValidQnums_kat():
    assert _t3().get_valid_qnums().flatten().tolist() == [1, 2, 4]


def test_ref_PutGetBlock_kat():
    T = _t3()
    blk = T.get_block(2)
    assert blk.shape == (3, 2) and blk.flatten().tolist() == [0.0] * 6
    T.put_block(np.arange(6.0).reshape(3, 2), 2)
    assert T.get_block(2).flatten().tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]


# ---- docstrings: Symmetry.py:14-27, :60-73 ; Bond.py:341-347 -----------------------
def test_ref_docstring_combine():
    b1 = orc.OBond(3, orc.BRA, qnums=[[0], [0], [1]])
    b2 = orc.OBond(4, orc.BRA, qnums=[[0], [2], [-3], [-1]])
    assert b1.qnums.flatten().tolist() == [1, 0, 0]
    assert b2.qnums.flatten().tolist() == [2, 0, -1, -3]
    assert orc.combine_qnums(b1.qnums, b2.qnums, [orc.U1]).flatten().tolist() == \
        [3, 1, 0, -2, 2, 0, -1, -3, 2, 0, -1, -3]
    z1 = orc.OBond(3, orc.BRA, qnums=[[0], [2], [1]], syms=[orc.Zn(4)])
    z2 = orc.OBond(4, orc.BRA, qnums=[[0], [2], [3], [1]], syms=[orc.Zn(4)])
    assert orc.combine_qnums(z1.qnums, z2.qnums, [orc.Zn(4)]).flatten().tolist() == \
        [1, 0, 3, 2, 0, 3, 2, 1, 3, 2, 1, 0]
    c = orc.OBond(2, orc.BRA, qnums=[[0, 1, -1], [1, 1, 0]])
    d = orc.OBond(2, orc.KET, qnums=[[1, 0, -1], [1, 0, 0]])
    assert orc.combine_qnums(c.qnums, d.qnums, [orc.U1] * 3).T.tolist() == \
        [[2, 2, 1, 1], [1, 1, 1, 1], [0, -1, -1, -2]]


# ---- SURVEY 8(a) vectors captured from the reference (sym_A of UniTensor.py:3603-3612) ----
def test_survey_sym_A_vectors():
    b1 = orc.OBond(3, orc.KET, qnums=[[0], [1], [2]])
    b2 = orc.OBond(4, orc.KET, qnums=[[-1], [2], [0], [2]])
    b3 = orc.OBond(5, orc.BRA, qnums=[[4], [2], [-1], [5], [1]])
    A = orc.OSym([b1, b2, b3], rowrank=2, labels=[10, 11, 12])
    assert A.block_qnums.flatten().tolist() == [-1, 1, 2, 4]
    assert A.bm.shapes == [(1, 1), (2, 1), (3, 1), (2, 1)]
    assert A.bm.ket_map.tolist() == [[3, 0], [3, 1], [2, 0], [1, 0], [-1, -1], [-1, -1], [1, 1], [-1, -1],
                                     [2, 1], [2, 2], [-1, -1], [0, 0]]
    assert A.bm.bra_map.tolist() == [[-1, -1], [3, 0], [2, 0], [1, 0], [0, 0]]
    assert [x.tolist() for x in A.bm.ket_inv] == [[[2, 3]], [[0, 3], [1, 2]], [[0, 2], [2, 0], [2, 1]], [[0, 0], [0, 1]]]
    v = 0
    for b in range(4):
        n = A.blocks[b].size
        A.blocks[b] = np.arange(v, v + n, dtype=np.float64).reshape(A.blocks[b].shape)
        v += n
    P = A.clone().permute([2, 0, 1], rowrank=1)
    assert P.mapper.tolist() == [2, 0, 1] and P.inv_mapper.tolist() == [1, 2, 0]
    assert P.block_qnums.flatten().tolist() == [-4, -2, -1, 1]
    P.contiguous_(vectorised=False)
    assert [b.tolist() for b in P.blocks] == [[[6, 7]], [[3, 4, 5]], [[1, 2]], [[0]]]
    Q = A.clone().permute([1, 0, 2], rowrank=2).contiguous_(vectorised=False)
    assert [b.tolist() for b in Q.blocks] == [[[0]], [[2], [1]], [[4], [5], [3]], [[6], [7]]]
    u1z2 = orc.unique_rows(np.array([[2, 1], [0, 1], [-2, 0], [1, 1], [0, 0], [0, 1]]))
    assert u1z2.tolist() == [[-2, 0], [0, 0], [0, 1], [1, 1], [2, 1]]


# ---- fixtures from the reference ------------------------------------------------------
def test_bond_goldens(golden):
    z, specs = golden("bond")
    for key, sp in specs.items():
        syms = [tuple(s) for s in sp["syms"]]
        a = orc.OBond(len(sp["raw_a"]), orc.KET, qnums=sp["raw_a"], syms=syms)
        b = orc.OBond(len(sp["raw_b"]), orc.KET, qnums=sp["raw_b"], syms=syms)
        assert np.array_equal(a.qnums, z[key + "/a_sorted"])
        assert np.array_equal(b.qnums, z[key + "/b_sorted"])
        assert np.array_equal(orc.unique_rows(a.qnums), z[key + "/a_unique"])
        assert np.array_equal(orc.comm_rows(orc.unique_rows(a.qnums), orc.unique_rows(b.qnums)), z[key + "/comm"])
        assert np.array_equal(orc.combine_qnums(a.qnums, b.qnums, syms), z[key + "/combined"])


def check_maps(z, key, bm):
    assert np.array_equal(bm.block_qnums, z[key + "/block_qnums"])
    assert np.array_equal(bm.ket_map, z[key + "/ket_map"])
    assert np.array_equal(bm.bra_map, z[key + "/bra_map"])
    assert np.array_equal(bm.accu_in, z[key + "/accu_in"])
    assert np.array_equal(bm.accu_out, z[key + "/accu_out"])
    assert np.array_equal(np.array(bm.shapes).reshape(-1, 2), z[key + "/shapes"])
    for b in range(len(bm.shapes)):
        assert np.array_equal(bm.ket_inv[b], z[key + "/ket_inv%d" % b])
        assert np.array_equal(bm.bra_inv[b], z[key + "/bra_inv%d" % b])


def test_blockmap_goldens(golden):
    z, specs = golden("blockmap")
    assert sum(1 for s in specs.values() if s["raises"]) >= 6
    for key, sp in specs.items():
        if sp["raises"]:
            with pytest.raises(TypeError):
                osym(sp)
            continue
        T = osym(sp)
        check_maps(z, key, T.bm)
        tin, tout = orc.total_qnums(T.bonds, T.braket, T.rowrank)
        assert np.array_equal(tin, z[key + "/tq_in"]) and np.array_equal(tout, z[key + "/tq_out"])


def test_relayout_goldens(golden):
    z, specs = golden("relayout")
    n_vec = 0
    for key, sp in specs.items():
        T = osym(sp, seed=sp["seed"])
        T.permute(sp["perm"], rowrank=sp["new_rowrank"])
        assert np.array_equal(T.block_qnums, z[key + "/logical_block_qnums"])
        assert np.array_equal(T.mapper, z[key + "/mapper"])
        assert np.array_equal(T.inv_mapper, z[key + "/inv_mapper"])
        assert bool(T.contiguous) == bool(z[key + "/contig_flag"])
        for s, q in enumerate(T.block_qnums):
            assert np.array_equal(T.get_block(*q.tolist()), z[key + "/getblk%d" % s])
        P = T.clone().contiguous_(vectorised=False)
        check_maps(z, key + "/out", P.bm)
        for b in range(int(z[key + "/out/nblk"])):
            assert np.array_equal(P.blocks[b], z[key + "/out/blk%d" % b])
        # the vectorised relayout (used for the big CPU baselines) must agree wherever every
        # element lands on a physical position (it asserts that itself)
        try:
            V = T.clone().contiguous_(vectorised=True)
        except AssertionError:
            V = None
        if V is not None:
            n_vec += 1
            for b in range(len(P.blocks)):
                assert np.array_equal(P.blocks[b], V.blocks[b])
        if len(T.block_qnums):
            q0 = T.block_qnums[0]
            blk = T.get_block(*q0.tolist())
            T.put_block(np.arange(blk.size, dtype=np.float64).reshape(blk.shape) + 100.0, *q0.tolist())
            for b in range(int(z[key + "/afterput/nblk"])):
                assert np.array_equal(T.blocks[b], z[key + "/afterput/blk%d" % b])
    assert n_vec >= 25


def test_contract_sym_goldens(golden):
    z, specs = golden("contract_sym")
    assert len(specs) >= 40
    for key, sp in specs.items():
        A = osym(sp["A"], seed=sp["seedA"])
        B = osym(sp["B"], seed=sp["seedB"])
        C = orc.contract_sym(A, B, vectorised=False)
        assert np.array_equal(C.labels, z[key + "/labels"])
        assert np.array_equal(C.braket, z[key + "/braket"])
        assert C.rowrank == int(z[key + "/rowrank"])
        check_maps(z, key + "/C", C.bm)
        for b in range(int(z[key + "/C/nblk"])):
            assert C.blocks[b].shape == z[key + "/C/blk%d" % b].shape
            assert rel_err(C.blocks[b], z[key + "/C/blk%d" % b]) <= 1e-13
        if key.startswith("cs"):
            C2 = orc.contract_sym(A, B, vectorised=True)
            for b in range(len(C.blocks)):
                assert np.array_equal(C.blocks[b], C2.blocks[b])


def test_contract_sym_equals_dense_einsum(golden):
    """Independent check: in the well-defined domain the block result
    equals tensordot of the densified operands."""
    z, specs = golden("contract_sym")
    for key, sp in specs.items():
        if not key.startswith("cs"):
            continue
        A = osym(sp["A"], seed=sp["seedA"])
        B = osym(sp["B"], seed=sp["seedB"])
        C = orc.contract_sym(A, B)
        same, a_ind, b_ind, a_free, b_free = orc._label_match(A.labels, B.labels)
        D = np.tensordot(A.to_dense(), B.to_dense(), (a_ind.tolist(), b_ind.tolist()))
        assert rel_err(C.to_dense(), D) <= 1e-13


def test_dense_goldens(golden):
    z, specs = golden("dense")
    for key, sp in specs.items():
        if not key.startswith("dc"):
            continue
        if sp["tagged"]:
            rng = np.random.Generator(np.random.PCG64(sp["seed"]))
            A = orc.ODense(rng.random(tuple(sp["A"]["dims"])), sp["A"]["labels"], sp["A"]["rowrank"], braket=sp["A"]["braket"])
            B = orc.ODense(rng.random(tuple(sp["B"]["dims"])), sp["B"]["labels"], sp["B"]["rowrank"], braket=sp["B"]["braket"])
        else:
            A, B = odense(sp["A"]), odense(sp["B"])
        C = orc.contract_dense(A, B)
        assert np.array_equal(C.labels, z[key + "/labels"])
        assert C.rowrank == int(z[key + "/rowrank"])
        assert C.data.shape == z[key + "/data"].shape
        assert rel_err(C.data, z[key + "/data"]) <= 1e-13
        if sp["tagged"]:
            assert np.array_equal(C.braket, z[key + "/braket"])


def _mats():
    out = []
    for j, (m, k) in enumerate([(3, 4), (4, 5), (5, 6), (6, 2)]):
        out.append(odense({"dims": [m, k], "labels": [0, 1], "rowrank": 1, "seed": 600 + j}))
    dg = odense({"dims": [4, 4], "labels": [0, 1], "rowrank": 1, "seed": 610, "is_diag": True})
    return out, dg


def test_matmul_goldens(golden):
    z, _ = golden("dense")
    mats, dg = _mats()
    assert rel_err(orc.matmul_dense(mats[0], mats[1]).data, z["mm/ab"]) <= 1e-13
    assert rel_err(orc.chain_matmul_dense(*mats).data, z["mm/chain"]) <= 1e-13
    assert rel_err(orc.matmul_dense(mats[0], dg).data, z["mm/a_diag"]) <= 1e-13
    assert rel_err(orc.matmul_dense(dg, mats[1]).data, z["mm/diag_b"]) <= 1e-13
    dd = orc.matmul_dense(dg, dg)
    assert dd.is_diag == bool(z["mm/diag_diag_isdiag"]) and rel_err(dd.data, z["mm/diag_diag"]) <= 1e-13


def test_network_goldens(golden):
    z, specs = golden("dense")
    for key in ("net0", "net1", "net2"):
        sp = specs[key]
        N = orc.ONetwork()
        N.from_text(sp["text"])
        for nm, (dims, rr, seed) in sp["tensors"].items():
            N.put(nm, odense({"dims": dims, "labels": list(range(len(dims))), "rowrank": rr, "seed": seed}))
        out, tout = N.launch(orc.contract_dense)
        want_labels = tout[0].tolist() + tout[1].tolist()
        pos = [out.labels.tolist().index(l) for l in want_labels]            # Network.py:422-423
        data = np.transpose(out.data, pos)
        assert want_labels == z[key + "/labels"].tolist()
        assert len(tout[0]) == int(z[key + "/rowrank"])
        assert rel_err(data, z[key + "/data"]) <= 1e-13
        for nm in sp["tensors"]:                                             # labels restored :379-380
            assert N.instances[nm].labels.tolist() == list(range(len(sp["tensors"][nm][0])))


def test_gauss_u1_generator():
    qs, deg = orc.gauss_u1_degeneracies(4096, -20, 19, 6.0)
    assert len(qs) == 40 and deg.sum() == 4096 and deg.min() == 1 and deg.max() == 271


def test_itebd_oracle_trajectory(golden):
    """oracle/itebd_numpy.py follows the reference's iTEBD.py loop: 40 energies from the same initial state."""
    import json
    from oracle import itebd_numpy as it
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "itebd.npz"))
    sp = json.loads(str(z["specs"]))
    init = {k: z[k] for k in ("A", "B", "la", "lb")}
    r = it.run(sp["J"], sp["Hx"], sp["chi"], 0.0, max_steps=sp["nsteps"], init=init)
    assert np.abs(np.array(r["energies"]) - z["energies"]).max() <= 1e-12


def test_itebd_oracle_energy():
    """Converged energy of config 1 (reference run in the build container: E = -1.252868)."""
    from oracle import itebd_numpy as it
    r = it.run(1.0, 0.5, 16, 1e-9, max_steps=20000, seed=3)
    assert abs(r["E"] - (-1.2528)) < 5e-4
