"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle's element loops are too
slow there): relayout round trips are the identity bit for bit, the grouped GEMM equals an independent per-sector
cuBLAS product of the relaid-out blocks within 1e-12, Contract is linear, the block map equals the oracle's
(vectorised numpy), fp32 equals fp64 within 1e-5, dense Contract equals torch.tensordot.
Configs: C3a / C3b (U1, chi=4096, ~40 sectors) and C2 (dense rank-4, chi=1024, d=2)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import obond, psym, rel_err
from oracle import tor10_oracle as orc

pytestmark = pytest.mark.gpu
DEV = os.environ.get("T10_TEST_DEVICE", "cuda")
CHI = int(os.environ.get("T10_FULL_CHI", "4096"))
TOL64, TOL32 = 1e-12, 1e-5


@pytest.fixture(scope="module")
def T():
    import tor10_b200
    return tor10_b200


@pytest.fixture(scope="module")
def c3a(T):
    import bench
    A, B = bench.c3a_specs(CHI)
    return A, B, psym(T, A, seed=1003, device=DEV), psym(T, B, seed=1004, device=DEV)


def test_blockmap_full_size_vs_oracle(T, c3a):
    A, _, a, _ = c3a
    bm = orc.build_blockmap([obond(b) for b in A["bonds"]], np.array([b["btype"] for b in A["bonds"]]), A["rowrank"])
    assert np.array_equal(a._block_qnums, bm.block_qnums)
    assert np.array_equal(a._Ket_mapper_blks, bm.ket_map) and np.array_equal(a._Bra_mapper_blks, bm.bra_map)
    assert [tuple(x.shape) for x in a.Storage] == [tuple(s) for s in bm.shapes]


def _fresh_copy(T, a):
    t = T.UniTensor(bonds=list(a.bonds), labels=a.labels.tolist(), rowrank=a.rowrank, device=torch.device(DEV))
    for i, x in enumerate(a.Storage):
        t.Storage[i] = x.clone()
    return t


@pytest.mark.parametrize("perm,rr", [([0, 2, 1, 3], 2), ([0, 2, 3, 1], 2), ([2, 0, 3, 1], 2), ([3, 0, 2, 1], 2)])
def test_relayout_full_size_vs_oracle(T, c3a, perm, rr):
    """Permute -> Contiguous at chi=4096 against the oracle's vectorised restatement of UniTensor.py:2108-2129,
    bit for bit -- including permutations that are NOT involutions (a wrong-but-consistent index algebra cannot
    pass by applying itself twice) and changed row ranks."""
    from helpers import osym
    A, _, a, _ = c3a
    o = osym(A, seed=1003)
    try:
        o.permute(perm, rowrank=rr)
    except TypeError:
        pytest.skip("no common sector for this split")
    o.contiguous_(vectorised=True)
    t = _fresh_copy(T, a)
    t.Permute(perm, rowrank=rr)
    p = t.Contiguous()
    assert p.is_contiguous()
    assert np.array_equal(p._block_qnums, o.bm.block_qnums)
    assert np.array_equal(p._Ket_mapper_blks, o.bm.ket_map) and np.array_equal(p._Bra_mapper_blks, o.bm.bra_map)
    assert len(p.Storage) == len(o.blocks)
    for x, y in zip(p.Storage, o.blocks):
        assert tuple(x.shape) == y.shape
        assert np.array_equal(x.cpu().numpy(), y)


def test_contract_full_size_vs_oracle(T, c3a):
    """The headline contraction against the oracle on the same seeded operands: maps bit-exact, blocks <= 1e-12."""
    from helpers import osym
    A, B, a, b = c3a
    mm = lambda x, y: torch.matmul(torch.from_numpy(x), torch.from_numpy(y)).numpy()  # noqa: E731
    oc = orc.contract_sym(osym(A, seed=1003), osym(B, seed=1004), matmul=mm)
    c = T.Contract(a, b)
    assert np.array_equal(c._block_qnums, oc.bm.block_qnums)
    assert np.array_equal(c._Ket_mapper_blks, oc.bm.ket_map) and np.array_equal(c._Bra_mapper_blks, oc.bm.bra_map)
    assert c.labels.tolist() == oc.labels.tolist() and c.rowrank == oc.rowrank
    assert len(c.Storage) == len(oc.blocks) >= 40
    worst = max(rel_err(x.cpu().numpy(), y) for x, y in zip(c.Storage, oc.blocks))
    assert worst <= TOL64, worst


def _per_sector_reference(T, a, b):
    """Independent path: relayout with the library, then one torch.matmul (cuBLAS) per matched sector --
    the reference's own loop (UniTensor.py:3727-3741) on the same blocks."""
    plan = T._plan.get_contract_plan(a, b)
    c = T.Contract(a, b)
    ta = a
    if plan.rel_a is not None:
        ta = T.UniTensor(bonds=list(a.bonds), labels=a.labels.tolist(), rowrank=a.rowrank, device=torch.device(DEV))
        for i, x in enumerate(a.Storage):
            ta.Storage[i] = x
        same, a_ind, b_ind, a_free, b_free = T._plan.label_match(a.labels, b.labels)
        ta.Permute(np.append(a_free, a_ind), rowrank=len(a_free), by_label=False)
        ta = ta.Contiguous()
    assert plan.rel_b is None
    worst = 0.0
    for ob, ab, bb in plan.matched:
        want = torch.matmul(ta.Storage[ab], b.Storage[bb])
        got = c.Storage[ob]
        worst = max(worst, float((got - want).abs().max() / want.abs().max()))
    return c, worst, len(plan.matched)


def test_contract_full_size_vs_per_sector_cublas(T, c3a):
    _, _, a, b = c3a
    c, worst, nsec = _per_sector_reference(T, a, b)
    assert nsec >= 40 and worst <= TOL64, worst
    # linearity in the second operand
    b2 = psym(T, c3a[1], seed=2024, device=DEV)
    bsum = b + b2
    c2, csum = T.Contract(a, b2), T.Contract(a, bsum)
    for x, y, z in zip(c.Storage, c2.Storage, csum.Storage):
        assert float((x + y - z).abs().max()) <= TOL64 * float(z.abs().max())
    # replay is bit-reproducible
    again = T.Contract(a, b)
    for x, y in zip(c.Storage, again.Storage):
        assert torch.equal(x, y)


def test_contract_full_size_fp32_vs_fp64(T, c3a):
    A, B, a, b = c3a
    c64 = T.Contract(a, b)
    a32 = psym(T, A, seed=1003, device=DEV, dtype=torch.float32)
    b32 = psym(T, B, seed=1004, device=DEV, dtype=torch.float32)
    c32 = T.Contract(a32, b32)
    assert c32.dtype == torch.float32
    for x, y in zip(c32.Storage, c64.Storage):
        assert float((x.double() - y).abs().max() / y.abs().max()) <= TOL32


def test_c3b_environment_step_full_size(T):
    """MPS-MPO-MPS environment step L[(x, w); (a)] . A[(a); (s, b)]: no relayout, 40 sectors."""
    v = orc.gauss_u1_qnums(CHI, -20, 19, 6.0).tolist()
    u1 = [["U1", 0]]
    w = [[0], [0], [0], [2], [-2]]
    p = [[1], [-1]]
    Ls = {"bonds": [{"dim": CHI, "btype": -1, "qnums": v, "syms": u1}, {"dim": 5, "btype": -1, "qnums": w, "syms": u1},
                    {"dim": CHI, "btype": 1, "qnums": v, "syms": u1}], "rowrank": 2, "labels": [0, 1, 2]}
    As = {"bonds": [{"dim": CHI, "btype": -1, "qnums": v, "syms": u1}, {"dim": 2, "btype": 1, "qnums": p, "syms": u1},
                    {"dim": CHI, "btype": 1, "qnums": v, "syms": u1}], "rowrank": 1, "labels": [2, 3, 4]}
    try:
        L, A = psym(T, Ls, seed=7, device=DEV), psym(T, As, seed=8, device=DEV)
    except TypeError:
        pytest.skip("no common sector for this synthetic leg")
    plan = T._plan.get_contract_plan(L, A)
    assert plan.rel_a is None and plan.rel_b is None
    c = T.Contract(L, A)
    for ob, ab, bb in plan.matched:
        want = torch.matmul(L.Storage[ab], A.Storage[bb])
        assert float((c.Storage[ob] - want).abs().max()) <= TOL64 * float(want.abs().max())


def test_c2_dense_contract_full_size(T):
    chi, d = (1024, 2) if CHI >= 1024 else (CHI, 2)
    g = torch.Generator().manual_seed(3)
    x = torch.rand((chi, d, d, chi), generator=g, dtype=torch.float64).to(DEV)
    y = torch.rand((d, chi, d, chi), generator=g, dtype=torch.float64).to(DEV)
    t1 = T.UniTensor(bonds=[T.Bond(chi), T.Bond(d), T.Bond(d), T.Bond(chi)], labels=[0, 2, 1, 3], rowrank=2,
                     device=torch.device(DEV))
    t2 = T.UniTensor(bonds=[T.Bond(d), T.Bond(chi), T.Bond(d), T.Bond(chi)], labels=[2, 3, 4, 5], rowrank=2,
                     device=torch.device(DEV))
    t1.Storage, t2.Storage = x, y
    c = T.Contract(t1, t2)
    # labels 2, 3 are contracted: t1 axes (1, 3) with t2 axes (0, 1); result [0, 1, 4, 5]
    want = torch.tensordot(x, y, dims=([1, 3], [0, 1]))
    assert c.labels.tolist() == [0, 1, 4, 5]
    got = c.Storage if c.Storage.shape == want.shape else c.Storage.reshape(want.shape)
    assert float((got - want).abs().max() / want.abs().max()) <= TOL64
