"""GPU parity: the CUDA path (through the C-ABI) against the oracle and the committed reference
fixtures, on the same seeded inputs.

Bars (north star): block maps / indices / labels / layouts bit-exact; data movement (relayout,
GetBlock, PutBlock) bit-exact; fp64 GEMM results within 1e-12 relative, fp32 within 1e-5.
"""
import copy
import os

import numpy as np
import pytest
import torch

from helpers import obond, osym, pbond, psym, rel_err
from oracle import tor10_oracle as orc

pytestmark = pytest.mark.gpu
DEV = os.environ.get("T10_TEST_DEVICE", "cuda")
TOL64, TOL32 = 1e-12, 1e-5


@pytest.fixture(scope="module")
def T():
    import tor10_b200
    if DEV == "cuda":
        assert torch.cuda.is_available(), "gpu tests need a CUDA device; there is no CPU path"
    return tor10_b200


def blocks_np(t):
    return [b.detach().cpu().numpy() for b in t.Storage]


def check_maps(z, key, t):
    """product attribute names == reference attribute names (UniTensor.py:381-390)"""
    assert np.array_equal(t._block_qnums, z[key + "/block_qnums"])
    assert np.array_equal(t._Ket_mapper_blks, z[key + "/ket_map"])
    assert np.array_equal(t._Bra_mapper_blks, z[key + "/bra_map"])
    assert np.array_equal(t._accu_off_in, z[key + "/accu_in"])
    assert np.array_equal(t._accu_off_out, z[key + "/accu_out"])
    shapes = np.array([tuple(s.shape) for s in t.Storage], dtype=np.int64).reshape(-1, 2)
    assert np.array_equal(shapes, z[key + "/shapes"])
    for b in range(len(shapes)):
        assert np.array_equal(t._Ket_invmapper_blks[b], z[key + "/ket_inv%d" % b])
        assert np.array_equal(t._Bra_invmapper_blks[b], z[key + "/bra_inv%d" % b])


# ---- family 1: block map -------------------------------------------------------------
def test_ref_kats(T):
    """Known answers of the reference's own tests (tests/test_UniTensor.py:86-137)."""
    b1 = T.Bond(3, T.BD_KET, qnums=[[0, 2, 1, 0], [1, 1, -1, 1], [2, -1, 1, 0]])
    b2 = T.Bond(4, T.BD_KET, qnums=[[-1, 0, -1, 3], [0, 0, -1, 2], [1, 0, 1, 0], [2, -2, -1, 1]])
    b3 = T.Bond(2, T.BD_BRA, qnums=[[-4, 3, 0, -1], [1, 1, -2, 3]])
    ut = T.UniTensor(bonds=[b1, b2, b3], rowrank=2, device=torch.device(DEV))
    tin, _ = ut.GetTotalQnums()
    want = T.Bond(12, T.BD_KET, qnums=[[+4, -3, +0, +1], [+3, -1, +2, +0], [+2, -1, +0, +2], [+1, -1, +0, +3],
                                      [+3, -1, -2, +2], [+2, +1, +0, +1], [+1, +1, -2, +3], [+0, +1, -2, +4],
                                      [+2, +0, +0, +1], [+1, +2, +2, +0], [+0, +2, +0, +2], [-1, +2, +0, +3]])
    got = tin.qnums[np.lexsort(tin.qnums.T[::-1])[::-1]]
    assert got.tolist() == want.qnums.tolist()
    c1 = T.Bond(3, T.BD_KET, qnums=[[0], [1], [2]])
    c2 = T.Bond(4, T.BD_KET, qnums=[[-1], [2], [0], [2]])
    c3 = T.Bond(5, T.BD_BRA, qnums=[[4], [2], [2], [5], [1]])
    t3 = T.UniTensor(bonds=[c1, c2, c3], rowrank=2, labels=[10, 11, 12], device=torch.device(DEV))
    assert t3.GetValidQnums().flatten().tolist() == [1, 2, 4]
    blk = t3.GetBlock(2)
    assert tuple(blk.shape) == (3, 2) and blk.Storage.flatten().tolist() == [0.0] * 6
    blk.SetElem([0, 1, 2, 3, 4, 5])
    t3.PutBlock(blk, 2)
    assert t3.GetBlock(2).Storage.flatten().tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]


def test_blockmap_goldens(T, golden):
    z, specs = golden("blockmap")
    for key, sp in specs.items():
        if sp["raises"]:
            with pytest.raises(TypeError):
                psym(T, sp, device=DEV)
            continue
        t = psym(T, sp, device=DEV)
        check_maps(z, key, t)
        tin, tout = t.GetTotalQnums(physical=False)
        assert np.array_equal(tin.qnums, z[key + "/tq_in"]) and np.array_equal(tout.qnums, z[key + "/tq_out"])


def test_blockmap_vs_oracle_large(T):
    """config-3 sized legs: every integer map of the product equals the oracle's."""
    v = orc.gauss_u1_qnums(256, -8, 7, 2.5).tolist()
    p = [[1], [-1]]
    spec = {"bonds": [{"dim": 256, "btype": -1, "qnums": v, "syms": [["U1", 0]]},
                      {"dim": 2, "btype": 1, "qnums": p, "syms": [["U1", 0]]},
                      {"dim": 2, "btype": -1, "qnums": p, "syms": [["U1", 0]]},
                      {"dim": 256, "btype": 1, "qnums": v, "syms": [["U1", 0]]}],
            "rowrank": 2, "labels": [0, 2, 1, 3]}
    t, o = psym(T, spec, device=DEV), osym(spec)
    assert np.array_equal(t._block_qnums, o.bm.block_qnums)
    assert np.array_equal(t._Ket_mapper_blks, o.bm.ket_map) and np.array_equal(t._Bra_mapper_blks, o.bm.bra_map)
    for b in range(len(o.bm.shapes)):
        assert tuple(t.Storage[b].shape) == o.bm.shapes[b]
        assert np.array_equal(t._Ket_invmapper_blks[b], o.bm.ket_inv[b])
        assert np.array_equal(t._Bra_invmapper_blks[b], o.bm.bra_inv[b])


# ---- family 2: relayout ---------------------------------------------------------------
def test_relayout_goldens(T, golden):
    z, specs = golden("relayout")
    n_exact = 0
    for key, sp in specs.items():
        t = psym(T, sp, seed=sp["seed"], device=DEV)
        o = osym(sp, seed=sp["seed"])
        t.Permute(sp["perm"], rowrank=sp["new_rowrank"])
        o.permute(sp["perm"], rowrank=sp["new_rowrank"])
        assert np.array_equal(t._block_qnums, z[key + "/logical_block_qnums"])
        assert np.array_equal(t._mapper, z[key + "/mapper"])
        assert np.array_equal(t._inv_mapper, z[key + "/inv_mapper"])
        assert bool(t.is_contiguous()) == bool(z[key + "/contig_flag"])
        # cases where a source element has no destination sector make the reference write through a
        # negative index (quirk Q3); the oracle's vectorised relayout asserts exactly that condition
        try:
            o.clone().contiguous_(vectorised=True)
            well_defined = True
        except AssertionError:
            well_defined = False
        for s, q in enumerate(t._block_qnums):
            got = t.GetBlock(*q.tolist()).Storage.cpu().numpy()
            assert got.shape == z[key + "/getblk%d" % s].shape
            if well_defined:
                assert np.array_equal(got, z[key + "/getblk%d" % s])
        P = t.Contiguous()
        assert P.is_contiguous()
        check_maps(z, key + "/out", P)
        if well_defined:
            n_exact += 1
            for b in range(int(z[key + "/out/nblk"])):
                assert np.array_equal(blocks_np(P)[b], z[key + "/out/blk%d" % b])
            if len(t._block_qnums):
                q0 = t._block_qnums[0]
                blk = t.GetBlock(*q0.tolist())
                n = blk.Storage.numel()
                blk.Storage = (torch.arange(n, dtype=torch.float64, device=blk.Storage.device)
                               .reshape(blk.Storage.shape) + 100.0)
                t.PutBlock(blk, *q0.tolist())
                for b in range(int(z[key + "/afterput/nblk"])):
                    assert np.array_equal(blocks_np(t)[b], z[key + "/afterput/blk%d" % b])
    assert n_exact >= 25


def test_relayout_inplace_and_properties(T, golden):
    """Contiguous_ mutates and returns self; the element multiset is conserved."""
    z, specs = golden("relayout")
    sp = specs["rl000"]
    t = psym(T, sp, seed=1, device=DEV)
    before = np.sort(np.concatenate([b.ravel() for b in blocks_np(t)]))
    r = t.Permute(sp["perm"], rowrank=sp["new_rowrank"]).Contiguous_()
    assert r is t and t.is_contiguous()
    after = np.sort(np.concatenate([b.ravel() for b in blocks_np(t)]))
    if len(before) == len(after):
        assert np.array_equal(before, after)


def _c3_spec(chi, qlo, qhi, sigma):
    v = orc.gauss_u1_qnums(chi, qlo, qhi, sigma).tolist()
    p = [[1], [-1]]
    u1 = [["U1", 0]]
    A = {"bonds": [{"dim": chi, "btype": -1, "qnums": v, "syms": u1}, {"dim": 2, "btype": 1, "qnums": p, "syms": u1},
                   {"dim": 2, "btype": -1, "qnums": p, "syms": u1}, {"dim": chi, "btype": 1, "qnums": v, "syms": u1}],
         "rowrank": 2, "labels": [0, 2, 1, 3]}
    B = {"bonds": [{"dim": 2, "btype": -1, "qnums": p, "syms": u1}, {"dim": chi, "btype": -1, "qnums": v, "syms": u1},
                   {"dim": 2, "btype": 1, "qnums": p, "syms": u1}, {"dim": chi, "btype": 1, "qnums": v, "syms": u1}],
         "rowrank": 2, "labels": [2, 3, 4, 5]}
    return A, B


def test_relayout_vs_oracle_medium(T):
    """C3a-shaped operand at chi=128 (the reference's element loop needs seconds here): bit-exact."""
    A, _ = _c3_spec(128, -6, 5, 2.0)
    t, o = psym(T, A, seed=77, device=DEV), osym(A, seed=77)
    t.Permute([0, 2, 1, 3], rowrank=2)
    o.permute([0, 2, 1, 3], rowrank=2).contiguous_(vectorised=True)
    P = t.Contiguous()
    assert np.array_equal(P._block_qnums, o.bm.block_qnums)
    for b, blk in enumerate(blocks_np(P)):
        assert np.array_equal(blk, o.blocks[b])
    # round trip back to the original layout is the identity
    P.Permute([0, 2, 1, 3], rowrank=2)
    Q = P.Contiguous()
    t0 = psym(T, A, seed=77, device=DEV)
    for x, y in zip(blocks_np(Q), blocks_np(t0)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("mode", ["segments", "runs", "index", "blocks"])
def test_relayout_replay_modes(T, mode, monkeypatch):
    """The four replays of one relayout -- segments of the element index (one warp per contiguous piece: default
    when the pieces are long), its runs with bisection, the element index itself, the factored block kernel --
    move exactly the same bits (oracle: vectorised restatement of UniTensor.py:2108-2129).  Narrow charge range =
    long runs (degeneracies of 40-80)."""
    monkeypatch.setenv("TOR10_RELAYOUT_MODE", "auto" if mode == "segments" else "index")
    monkeypatch.setenv("TOR10_RELAYOUT_RUNS", "1" if mode == "runs" else "auto")
    monkeypatch.setenv("TOR10_RELAYOUT_INDEX", "0" if mode == "blocks" else "1")
    T._plan.clear_caches()
    A, B = _c3_spec(256, -2, 2, 1.2)
    t, o = psym(T, A, seed=31, device=DEV), osym(A, seed=31)
    t.Permute([0, 2, 1, 3], rowrank=2)
    o.permute([0, 2, 1, 3], rowrank=2).contiguous_(vectorised=True)
    P = t.Contiguous()
    for b, blk in enumerate(blocks_np(P)):
        assert np.array_equal(blk, o.blocks[b])
    # the same relayout inside Contract (padded operand scratch, sector subset = all)
    a, b = psym(T, A, seed=31, device=DEV), psym(T, B, seed=32, device=DEV)
    Cc = T.Contract(a, b)
    pl = T._plan.get_contract_plan(a, b)
    assert (pl.rel_a.d_seg is not None) == (mode == "segments")
    assert (pl.rel_a.d_runs is not None) == (mode == "runs")
    assert (pl.rel_a.d_index is not None) == (mode == "index")
    oc = orc.contract_sym(osym(A, seed=31), osym(B, seed=32))
    for blk, ref in zip(blocks_np(Cc), oc.blocks):
        assert rel_err(blk, ref) <= TOL64
    T._plan.clear_caches()


# ---- family 3 + Contract --------------------------------------------------------------
def test_contract_sym_goldens(T, golden):
    z, specs = golden("contract_sym")
    for key, sp in specs.items():
        A = psym(T, sp["A"], seed=sp["seedA"], device=DEV)
        B = psym(T, sp["B"], seed=sp["seedB"], device=DEV)
        a_before = [x.copy() for x in blocks_np(A)]
        Cc = T.Contract(A, B)
        assert np.array_equal(Cc.labels, z[key + "/labels"])
        assert np.array_equal(Cc.braket, z[key + "/braket"])
        assert Cc.rowrank == int(z[key + "/rowrank"])
        check_maps(z, key + "/C", Cc)
        for b, blk in enumerate(blocks_np(Cc)):
            assert blk.shape == z[key + "/C/blk%d" % b].shape
            assert rel_err(blk, z[key + "/C/blk%d" % b]) <= TOL64
        # operands untouched (the reference deep-copies them, UniTensor.py:3703-3704)
        assert A.labels.tolist() == sp["A"]["labels"] and A.rowrank == sp["A"]["rowrank"]
        for x, y in zip(blocks_np(A), a_before):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("cfg", ["0", "1", "2", "3", "8", "40", "42", "46", "100"])
def test_contract_c3a_small_all_tile_shapes(T, cfg, monkeypatch):
    """C3a at chi=192 with every GEMM tile configuration (+ the SIMT cross-check kernel) vs the oracle."""
    monkeypatch.setenv("TOR10_GGEMM_CFG", cfg)
    T._plan.clear_caches()
    A, B = _c3_spec(192, -7, 6, 2.5)
    a, b = psym(T, A, seed=1003, device=DEV), psym(T, B, seed=1004, device=DEV)
    oa, ob = osym(A, seed=1003), osym(B, seed=1004)
    Cc = T.Contract(a, b)
    oc = orc.contract_sym(oa, ob)
    assert np.array_equal(Cc._block_qnums, oc.bm.block_qnums)
    assert Cc.labels.tolist() == oc.labels.tolist() and Cc.rowrank == oc.rowrank
    for blk, ref in zip(blocks_np(Cc), oc.blocks):
        assert blk.shape == ref.shape and rel_err(blk, ref) <= TOL64
    T._plan.clear_caches()


def test_contract_fp32(T):
    A, B = _c3_spec(96, -5, 4, 2.0)
    a = psym(T, A, seed=5, device=DEV, dtype=torch.float32)
    b = psym(T, B, seed=6, device=DEV, dtype=torch.float32)
    oa, ob = osym(A, seed=5), osym(B, seed=6)
    Cc = T.Contract(a, b)
    oc = orc.contract_sym(oa, ob)
    assert Cc.dtype == torch.float32
    for blk, ref in zip(blocks_np(Cc), oc.blocks):
        assert rel_err(blk, ref) <= TOL32


@pytest.mark.parametrize("f32cfg", ["50", "0"])
@pytest.mark.parametrize("chi", [40, 96, 200])
def test_contract_fp32_kernels(T, chi, f32cfg, monkeypatch):
    """fp32 grouped GEMM, both kernels: 50 = tcgen05 / TMEM (kind::tf32, three-pass split; the default) and
    0 = SIMT.  Within the north star's 1e-5 of the fp64 oracle -- a single TF32 pass would sit at ~5e-4."""
    import ctypes
    monkeypatch.setenv("TOR10_F32_CFG", f32cfg)
    T._plan.clear_caches()
    A, B = _c3_spec(chi, -5, 4, 2.0)
    a = psym(T, A, seed=15, device=DEV, dtype=torch.float32)
    b = psym(T, B, seed=16, device=DEV, dtype=torch.float32)
    oa, ob = osym(A, seed=15), osym(B, seed=16)
    Cc = T.Contract(a, b)
    if DEV == "cuda":
        torch.cuda.synchronize()
        flag = ctypes.c_int(0)
        T._lib.check(T._lib.load().t10_ggemm_tf32_status(ctypes.byref(flag)))
        assert flag.value == 0, "a tcgen05 CTA gave up waiting on its MMA barrier"
    oc = orc.contract_sym(oa, ob)
    assert Cc.dtype == torch.float32
    worst = max(rel_err(blk, ref) for blk, ref in zip(blocks_np(Cc), oc.blocks))
    assert worst <= TOL32, worst
    # dense fp32 Matmul with ragged sizes through the same kernel
    g = torch.Generator().manual_seed(chi)
    x = torch.rand((chi + 3, 77), generator=g, dtype=torch.float32).to(DEV)
    y = torch.rand((77, 2 * chi + 1), generator=g, dtype=torch.float32).to(DEV)
    ux = T.UniTensor(bonds=[T.Bond(chi + 3), T.Bond(77)], rowrank=1, device=torch.device(DEV), dtype=torch.float32)
    uy = T.UniTensor(bonds=[T.Bond(77), T.Bond(2 * chi + 1)], rowrank=1, device=torch.device(DEV), dtype=torch.float32)
    ux.Storage, uy.Storage = x, y
    got = T.Matmul(ux, uy).Storage.cpu().numpy()
    want = x.cpu().double().numpy() @ y.cpu().double().numpy()
    assert rel_err(got, want) <= TOL32
    T._plan.clear_caches()


def test_matmul_fp32_tcgen05_long_k(T):
    """Default fp32 kernel on a long, ragged k loop (1000 = 31 slabs + 8: four accumulator chunks, the last
    one partial) with more tiles than SMs (TMEM reused across tiles) and all-positive data, the worst case for
    the tensor core's truncating accumulator: the chunked round-to-nearest drain must keep fp32 quality."""
    import ctypes
    g = torch.Generator().manual_seed(7)
    x = torch.rand((1601, 1000), generator=g, dtype=torch.float32).to(DEV)
    y = torch.rand((1000, 1702), generator=g, dtype=torch.float32).to(DEV)
    ux = T.UniTensor(bonds=[T.Bond(1601), T.Bond(1000)], rowrank=1, device=torch.device(DEV), dtype=torch.float32)
    uy = T.UniTensor(bonds=[T.Bond(1000), T.Bond(1702)], rowrank=1, device=torch.device(DEV), dtype=torch.float32)
    ux.Storage, uy.Storage = x, y
    got = T.Matmul(ux, uy).Storage.cpu().numpy()
    if DEV == "cuda":
        flag = ctypes.c_int(0)
        T._lib.check(T._lib.load().t10_ggemm_tf32_status(ctypes.byref(flag)))
        assert flag.value == 0
    want = x.cpu().double().numpy() @ y.cpu().double().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err <= 3e-6, err      # measured 1.9e-6; cuBLAS fp32 sits at ~4e-6 on such data


@pytest.mark.parametrize("shape", [(1, 4096, 1), (3, 5000, 70), (64, 1024, 64), (130, 2051, 100)])
def test_dense_contract_split_k(T, shape):
    """Long-K products with a handful of output tiles (iTEBD norms / expectation values) run as K-slices of one
    grouped launch + an ordered reduction: same result as the fp64 reference within 1e-12, and bit-reproducible."""
    M, K, N = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.rand((M, K), generator=g, dtype=torch.float64).to(DEV)
    y = torch.rand((K, N), generator=g, dtype=torch.float64).to(DEV)
    ux = T.UniTensor(bonds=[T.Bond(M), T.Bond(K)], rowrank=1, device=torch.device(DEV))
    uy = T.UniTensor(bonds=[T.Bond(K), T.Bond(N)], rowrank=1, device=torch.device(DEV))
    ux.Storage, uy.Storage = x, y
    got = T.Matmul(ux, uy).Storage
    again = T.Matmul(ux, uy).Storage
    assert torch.equal(got, again)
    want = x.cpu().numpy() @ y.cpu().numpy()
    assert rel_err(got.cpu().numpy(), want) <= TOL64


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("chi", [640, 1000])
def test_contract_stream_k_forced(T, chi, tma, monkeypatch):
    """Stream-K forced on small groups: shares of a few k-steps, so tiles are cut into MANY pieces (several parked
    partials per tile, added in slot order).  Same result as the oracle within 1e-12 and bit-reproducible -- on the
    TMA-fed kernel (ggemm_tma.cu, default) and on the cp.async kernel it replaces (TOR10_GGEMM_TMA=0)."""
    monkeypatch.setenv("TOR10_GGEMM_CFG", "46")
    monkeypatch.setenv("TOR10_GGEMM_STREAMK", "force")
    monkeypatch.setenv("TOR10_GGEMM_TMA", tma)
    T._plan.clear_caches()
    A, B = _c3_spec(chi, -3, 3, 1.5)
    a, b = psym(T, A, seed=21, device=DEV), psym(T, B, seed=22, device=DEV)
    c1 = T.Contract(a, b)
    plan = T._plan.get_contract_plan(a, b)
    assert plan.gemm.tile_cfg == 46 and plan.gemm.sk_parked > 0
    assert plan.gemm.tma == (tma == "1") and plan.gemm.streamk == (tma == "0")
    c2 = T.Contract(a, b)
    oc = orc.contract_sym(osym(A, seed=21), osym(B, seed=22))
    for x, y, ref in zip(blocks_np(c1), blocks_np(c2), oc.blocks):
        assert np.array_equal(x, y)
        assert rel_err(x, ref) <= TOL64
    if DEV == "cuda":
        assert int(plan.gemm.d_sk_flags[plan.gemm.sk_nslots].item()) == 0, "a parked partial never arrived"
    T._plan.clear_caches()


def test_contract_plan_is_cached_and_repeatable(T):
    A, B = _c3_spec(64, -4, 3, 1.5)
    a, b = psym(T, A, seed=11, device=DEV), psym(T, B, seed=12, device=DEV)
    c1 = T.Contract(a, b)
    n = dict(T._plan.STATS)
    c2 = T.Contract(a, b)
    assert T._plan.STATS == n, "second identical Contract must not rebuild any plan"
    for x, y in zip(blocks_np(c1), blocks_np(c2)):
        assert np.array_equal(x, y)


def test_contract_equals_dense_einsum(T):
    """Independent of the oracle: densify the product's result through its own inverse maps and compare
    with a dense tensordot of the densified operands."""
    A, B = _c3_spec(48, -3, 3, 1.2)
    a, b = psym(T, A, seed=21, device=DEV), psym(T, B, seed=22, device=DEV)

    def densify(t):
        assert t.is_contiguous()
        dims = [bd.dim for bd in t.bonds]
        D = np.zeros(dims)
        for blk, ki, bi in zip(blocks_np(t), t._Ket_invmapper_blks, t._Bra_invmapper_blks):
            for i in range(blk.shape[0]):
                for j in range(blk.shape[1]):
                    D[tuple(ki[i]) + tuple(bi[j])] = blk[i, j]
        return D

    Cc = T.Contract(a, b)
    Da, Db = densify(a), densify(b)
    # a labels [0,2,1,3], b labels [2,3,4,5]: contract a.pos1<->b.pos0 (label 2), a.pos3<->b.pos1 (label 3)
    want = np.tensordot(Da, Db, axes=([1, 3], [0, 1]))
    assert rel_err(densify(Cc), want) <= TOL64


def test_contract_errors(T):
    A, B = _c3_spec(8, -1, 1, 1.0)
    a, b = psym(T, A, device=DEV), psym(T, B, device=DEV)
    d = T.UniTensor(bonds=[T.Bond(2), T.Bond(2)], rowrank=1, device=torch.device(DEV))
    with pytest.raises(TypeError):
        T.Contract(a, d)                                   # UniTensor.py:3682-3683
    b2 = copy.deepcopy(b)
    b2.SetLabels([7, 8, 9, 10])
    with pytest.raises(Exception):
        T.Contract(a, b2)                                  # symmetric outer product: "Developing" :3745-3747
    a2 = copy.deepcopy(a)
    a2.SetLabels([2, 0, 1, 3])                             # now a KET bond would meet b's KET bond
    with pytest.raises(Exception):
        T.Contract(a2, b)                                  # :3700-3701


def test_storage_assignment_is_packed_back(T):
    """User code assigns Storage[b] (reference PutBlock does, UniTensor.py:2993); the arena follows."""
    A, B = _c3_spec(16, -2, 2, 1.0)
    a, b = psym(T, A, seed=1, device=DEV), psym(T, B, seed=2, device=DEV)
    a.Storage[0] = torch.full_like(a.Storage[0], 3.0)
    oa, ob = osym(A, seed=1), osym(B, seed=2)
    oa.blocks[0][:] = 3.0
    Cc = T.Contract(a, b)
    oc = orc.contract_sym(oa, ob)
    for blk, ref in zip(blocks_np(Cc), oc.blocks):
        assert rel_err(blk, ref) <= TOL64


# ---- dense path, Matmul, Network ----------------------------------------------------------
def _pdense(T, d, tagged_braket=None):
    rng = np.random.Generator(np.random.PCG64(d["seed"]))
    shape = (d["dims"][0],) if d.get("is_diag") else tuple(d["dims"])
    data = torch.from_numpy(rng.random(shape)).to(DEV)
    if tagged_braket is None:
        bonds = [T.Bond(int(x)) for x in d["dims"]]
    else:
        bonds = [T.Bond(int(x), T.BD_BRA if t == 1 else T.BD_KET) for x, t in zip(d["dims"], tagged_braket)]
    t = T.UniTensor(bonds=bonds, rowrank=d["rowrank"], labels=d["labels"], is_diag=bool(d.get("is_diag")),
                    device=torch.device(DEV))
    t.Storage = data
    return t


def test_dense_contract_goldens(T, golden):
    z, specs = golden("dense")
    for key, sp in specs.items():
        if not key.startswith("dc"):
            continue
        if sp["tagged"]:
            rng = np.random.Generator(np.random.PCG64(sp["seed"]))
            A = _pdense(T, dict(sp["A"], seed=0), sp["A"]["braket"])
            B = _pdense(T, dict(sp["B"], seed=0), sp["B"]["braket"])
            A.Storage = torch.from_numpy(rng.random(tuple(sp["A"]["dims"]))).to(DEV)
            B.Storage = torch.from_numpy(rng.random(tuple(sp["B"]["dims"]))).to(DEV)
        else:
            A, B = _pdense(T, sp["A"]), _pdense(T, sp["B"])
        Cc = T.Contract(A, B)
        assert np.array_equal(Cc.labels, z[key + "/labels"])
        assert Cc.rowrank == int(z[key + "/rowrank"])
        got = Cc.Storage.contiguous().cpu().numpy()
        assert got.shape == z[key + "/data"].shape
        assert rel_err(got, z[key + "/data"]) <= TOL64
        if sp["tagged"]:
            assert np.array_equal(Cc.braket, z[key + "/braket"])


def test_dense_permute_kernel(T):
    """Contiguous() of permuted dense tensors == torch's own strided copy, bit-exact."""
    g = torch.Generator().manual_seed(3)
    for shape, perm in [((7, 5, 3), (2, 0, 1)), ((64, 2, 2, 64), (0, 2, 1, 3)), ((33, 65), (1, 0)),
                        ((4, 6, 5, 3, 2), (4, 2, 0, 3, 1)), ((128, 96), (1, 0)), ((5,), (0,)),
                        ((2, 3, 4), (0, 1, 2)), ((16, 1, 8), (2, 1, 0))]:
        for dt in (torch.float64, torch.float32):
            x = torch.rand(shape, generator=g, dtype=dt).to(DEV)
            t = T.UniTensor(bonds=[T.Bond(int(s)) for s in shape], rowrank=1, device=torch.device(DEV), dtype=dt)
            t.Storage = x
            t.Permute(list(perm))
            c = t.Contiguous()
            assert c.is_contiguous()
            assert torch.equal(c.Storage, x.permute(*perm).contiguous())


def test_dense_permute_tma_path(T, monkeypatch):
    """Permutes that keep the innermost bond go through the TMA load/store ring (UTMALDG/UTMASTG); ragged
    extents exercise the hardware edge clipping.  Bit-exact against torch's strided copy."""
    monkeypatch.setenv("TOR10_PERMUTE_TMA_MIN_BYTES", "0")
    g = torch.Generator().manual_seed(5)
    cases = [((1024, 2, 2, 1024), (0, 2, 1, 3)), ((37, 5, 3, 130), (2, 0, 1, 3)), ((300, 7, 258), (1, 0, 2)),
             ((9, 4, 6, 5, 64), (3, 1, 0, 2, 4)), ((513, 3, 1026), (1, 0, 2)), ((2, 3, 4, 5, 6, 8), (4, 2, 0, 3, 1, 5))]
    for shape, perm in cases:
        for dt in (torch.float64, torch.float32):
            x = torch.rand(shape, generator=g, dtype=dt).to(DEV)
            got = T.UniTensor._contiguous_dense(x.permute(*perm))
            assert got.is_contiguous() and torch.equal(got, x.permute(*perm).contiguous())
    # sliced (non-zero offset, strided) source
    x = torch.rand((64, 6, 512), generator=g, dtype=torch.float64).to(DEV)
    v = x[:, 1:5, 128:384].permute(1, 0, 2)
    assert torch.equal(T.UniTensor._contiguous_dense(v), v.contiguous())


def test_matmul_goldens(T, golden):
    z, _ = golden("dense")
    mats = [_pdense(T, {"dims": [m, k], "labels": [0, 1], "rowrank": 1, "seed": 600 + j})
            for j, (m, k) in enumerate([(3, 4), (4, 5), (5, 6), (6, 2)])]
    dg = _pdense(T, {"dims": [4, 4], "labels": [0, 1], "rowrank": 1, "seed": 610, "is_diag": True})
    assert rel_err(T.Matmul(mats[0], mats[1]).Storage.cpu().numpy(), z["mm/ab"]) <= TOL64
    assert rel_err(T.Chain_matmul(*mats).Storage.cpu().numpy(), z["mm/chain"]) <= TOL64
    assert rel_err(T.Matmul(mats[0], dg).Storage.cpu().numpy(), z["mm/a_diag"]) <= TOL64
    assert rel_err(T.Matmul(dg, mats[1]).Storage.cpu().numpy(), z["mm/diag_b"]) <= TOL64
    # diag x diag: the reference returns the dot product of the diagonals (bug); the product
    # of two diagonal matrices is returned here.  Pin the documented deviation.
    dd = T.Matmul(dg, dg)
    assert dd.is_diag and rel_err(dd.Storage.cpu().numpy(), dg.Storage.cpu().numpy() ** 2) <= TOL64
    assert rel_err(dd.Storage.sum().item(), z["mm/diag_diag"]) <= TOL64


def test_network_goldens(T, golden, tmp_path):
    z, specs = golden("dense")
    for key in ("net0", "net1", "net2"):
        sp = specs[key]
        f = tmp_path / (key + ".net")
        f.write_text(sp["text"])
        N = T.Network()
        N.Fromfile(str(f))
        for nm, (dims, rr, seed) in sp["tensors"].items():
            N.Put(nm, _pdense(T, {"dims": dims, "labels": list(range(len(dims))), "rowrank": rr, "seed": seed}))
        R = N.Launch()
        assert R.labels.tolist() == z[key + "/labels"].tolist()
        assert R.rowrank == int(z[key + "/rowrank"])
        assert rel_err(R.Storage.contiguous().cpu().numpy(), z[key + "/data"]) <= TOL64
        for nm in sp["tensors"]:
            assert N.instances[nm].labels.tolist() == list(range(len(sp["tensors"][nm][0])))


def test_network_symmetric_extension(T):
    """Network over symmetric tensors (refused by the reference, Network.py:301-302) == the same chain of
    Contract calls."""
    A, B = _c3_spec(24, -2, 2, 1.0)
    a, b = psym(T, A, seed=31, device=DEV), psym(T, B, seed=32, device=DEV)
    N = T.Network()
    N.Fromstring("A: 0 2; 1 3\nB: 2 3; 4 5\nTOUT: 0 1; 4 5\nOrder: A,B\n")
    N.Put("A", a)
    N.Put("B", b)
    R = N.Launch()
    a2, b2 = copy.deepcopy(a), copy.deepcopy(b)
    want = T.Contract(b2, a2)          # evaluator pops the most recent operand first
    want.Permute([0, 1, 4, 5], 2, by_label=True)
    assert R.labels.tolist() == want.labels.tolist() == [0, 1, 4, 5]
    Rc, Wc = R.Contiguous(), want.Contiguous()
    for x, y in zip(blocks_np(Rc), blocks_np(Wc)):
        assert np.array_equal(x, y)
    oc = orc.contract_sym(osym(A, seed=31), osym(B, seed=32))
    for x, y in zip(blocks_np(Rc), oc.blocks):
        assert rel_err(x, y) <= TOL64


# ---- config 1: iTEBD end to end ---------------------------------------------------------------
def test_itebd_trajectory_matches_reference(T):
    """examples/itebd.py (10 Contract + Svd_truncate + Inverse per step on the GPU kernels) from the seeded
    initial state of tests/golden/itebd.npz: energies == the unmodified reference's, step by step."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(root, "examples", "itebd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    z = np.load(os.path.join(root, "tests", "golden", "itebd.npz"))
    sp = json.loads(str(z["specs"]))
    init = {k: z[k] for k in ("A", "B", "la", "lb")}
    out = ex.run(sp["J"], sp["Hx"], sp["chi"], 0.0, max_steps=sp["nsteps"], init=init)
    got = np.array(out["energies"])
    assert got.shape == z["energies"].shape
    assert np.abs(got - z["energies"]).max() <= 1e-9


def test_itebd_graphed_matches_eager_and_reference(T):
    """The same loop with the step captured into one CUDA graph (tor10_b200.graph.StaticLoop, examples/itebd.py
    run_graphed): energies equal the eager run's and the unmodified reference's, step by step."""
    if DEV != "cuda":
        pytest.skip("CUDA graphs need the device")
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("itebd_example", os.path.join(root, "examples", "itebd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    z = np.load(os.path.join(root, "tests", "golden", "itebd.npz"))
    sp = json.loads(str(z["specs"]))
    init = {k: z[k] for k in ("A", "B", "la", "lb")}
    eager = np.array(ex.run(sp["J"], sp["Hx"], sp["chi"], 0.0, max_steps=sp["nsteps"], init=init)["energies"])
    out = ex.run_graphed(sp["J"], sp["Hx"], sp["chi"], 0.0, max_steps=sp["nsteps"], init=init)
    got = np.array(out["energies"])
    assert out["steps"] == sp["nsteps"] and got.shape == z["energies"].shape
    assert np.abs(got - eager).max() <= 1e-12
    assert np.abs(got - z["energies"]).max() <= 1e-9


def test_permute_without_common_sector_keeps_an_empty_table(T):
    """The reference's Permute only recomputes `_block_qnums` (UniTensor.py:2310-2313): when the new row and column
    spaces share no charge the table is EMPTY and nothing raises until the tensor is relaid out (found by
    tests/fuzz_vs_reference.py; Z2 with a lone BRA bond in the row space, quirk Q3)."""
    z2 = [["Zn", 2]]
    spec = {"bonds": [{"dim": 1, "btype": -1, "qnums": [[1]], "syms": z2}, {"dim": 1, "btype": -1, "qnums": [[0]], "syms": z2},
                      {"dim": 2, "btype": 1, "qnums": [[1], [0]], "syms": z2}], "rowrank": 2, "labels": [8, 11, 6]}
    t = psym(T, spec, seed=3, device=DEV)
    t.Permute([2, 1, 0], rowrank=2)
    q = t.GetValidQnums()
    assert np.asarray(q).shape == (0, 1) and np.asarray(t._block_qnums).dtype == np.int64
    with pytest.raises(TypeError):
        t.Contiguous()


def test_dense_reshape_view_combine_follow_the_reference(T):
    """Glue either side of the path, pinned to the reference's behaviour (found by tests/fuzz_vs_reference.py ops):
    View refuses a non-contiguous tensor and returns a COPY (UniTensor.py:2583-2594); Reshape behaves like
    torch.reshape on the strided storage (:2404-2406); dense CombineBonds without permute_back leaves the fused
    bond alone in its space: rowrank 1 when it goes first, rank-1 when it goes last (:4126-4139)."""
    dev = torch.device(DEV)
    x = T.UniTensor(bonds=[T.Bond(3), T.Bond(2), T.Bond(4)], rowrank=1, labels=[5, 6, 7], device=dev)
    x.Storage = torch.arange(24, dtype=torch.float64).reshape(3, 2, 4).to(dev)
    v = x.View([6, 4], 1)
    v.Storage[0, 0] = -1.0
    assert float(x.Storage[0, 0, 0]) == 0.0                      # not an alias
    x.Permute([1, 0, 2])
    assert not x.is_contiguous()
    with pytest.raises(Exception):
        x.View([6, 4], 1)
    same_shape = x.Reshape([2, 3, 4], 1)                          # viewable: strides carry over
    assert not same_shape.is_contiguous()
    flat = x.Reshape([6, 4], 1)                                   # needs a copy
    assert flat.is_contiguous()
    want = torch.arange(24, dtype=torch.float64).reshape(3, 2, 4).permute(1, 0, 2).reshape(6, 4)
    assert torch.equal(flat.Storage.cpu(), want)
    y = T.UniTensor(bonds=[T.Bond(2), T.Bond(3), T.Bond(4), T.Bond(5)], rowrank=3, labels=[1, 2, 3, 4], device=dev)
    y.Storage = torch.arange(120, dtype=torch.float64).reshape(2, 3, 4, 5).to(dev)
    a = copy.deepcopy(y)
    a.CombineBonds([3, 1])                                        # first constituent in row space -> fused bond first
    assert a.labels.tolist() == [3, 2, 4] and a.rowrank == 1 and list(a.shape) == [8, 3, 5]
    assert torch.equal(a.Storage.cpu(), torch.arange(120, dtype=torch.float64).reshape(2, 3, 4, 5).permute(2, 0, 1, 3).reshape(8, 3, 5))
    b = copy.deepcopy(y)
    b.CombineBonds([4, 2])                                        # first constituent in column space -> fused bond last
    assert b.labels.tolist() == [1, 3, 4] and b.rowrank == 2 and list(b.shape) == [2, 4, 15]


def test_whole_transpose_follows_the_reference(T):
    """UniTensor.py:1386-1423 (found by tests/fuzz_vs_reference.py symops): a transposed COPY -- row and column spaces
    swapped, bra/ket tags exchanged on tagged tensors, `self` untouched; on a symmetric tensor the relaid-out result
    holds the transposed blocks under the same sector table."""
    dev = torch.device(DEV)
    x = T.UniTensor(bonds=[T.Bond(2), T.Bond(3), T.Bond(4)], rowrank=1, labels=[5, 6, 7], device=dev)
    x.Storage = torch.arange(24, dtype=torch.float64).reshape(2, 3, 4).to(dev)
    y = x.Whole_transpose()
    assert x.labels.tolist() == [5, 6, 7] and x.rowrank == 1
    assert y.labels.tolist() == [6, 7, 5] and y.rowrank == 2 and list(y.shape) == [3, 4, 2]
    assert torch.equal(y.Storage.contiguous().cpu(), torch.arange(24, dtype=torch.float64).reshape(2, 3, 4).permute(1, 2, 0))
    t = T.UniTensor(bonds=[T.Bond(2, T.BD_KET), T.Bond(3, T.BD_BRA)], rowrank=1, labels=[1, 2], device=dev)
    u = t.Whole_transpose()
    assert [b.bondType for b in t.bonds] == [T.BD_KET, T.BD_BRA]
    assert u.labels.tolist() == [2, 1] and [b.bondType for b in u.bonds] == [T.BD_KET, T.BD_BRA] and u.braket.tolist() == [-1, 1]
    u1 = [T.Symmetry.U1()]
    s = T.UniTensor(bonds=[T.Bond(3, T.BD_KET, qnums=[[1], [0], [-1]], sym_types=u1),
                           T.Bond(2, T.BD_KET, qnums=[[1], [0]], sym_types=u1),
                           T.Bond(4, T.BD_BRA, qnums=[[2], [1], [0], [-1]], sym_types=u1)],
                    rowrank=2, labels=[1, 2, 3], device=dev)
    for i in range(len(s.Storage)):
        s.Storage[i] = (torch.arange(s.Storage[i].numel(), dtype=torch.float64).reshape(s.Storage[i].shape) + 10 * i).to(dev)
    before = [b.clone() for b in s.Storage]
    tr = s.Whole_transpose()
    assert not tr.is_contiguous()
    w = tr.Contiguous()
    assert w.labels.tolist() == [3, 1, 2] and w.rowrank == 1 and w.braket.tolist() == [-1, 1, 1]
    assert np.array_equal(w._block_qnums, s._block_qnums)
    for a, b in zip(before, w.Storage):
        assert torch.equal(a.t().contiguous(), b)
    for a, b in zip(before, s.Storage):
        assert torch.equal(a, b)


def test_contract_of_transposed_operands(T):
    """(A.B)^T = B^T.A^T through Whole_transpose: the operands carry exchanged bra/ket tags, so their stored sector
    tables no longer describe them even where the bond order is back to the stored one -- Contract must relayout
    instead of reading them in place (the reference keeps stale blocks in that state)."""
    A, B = _c3_spec(48, -3, 3, 1.5)
    a, b = psym(T, A, seed=61, device=DEV), psym(T, B, seed=62, device=DEV)
    c = T.Contract(a, b)
    ct = T.Contract(b.Whole_transpose(), a.Whole_transpose())
    want = c.Whole_transpose()
    want.Permute([int(x) for x in ct.labels], rowrank=int(ct.rowrank), by_label=True)
    want = want.Contiguous()
    got = ct if ct.is_contiguous() else ct.Contiguous()
    assert np.array_equal(got._block_qnums, want._block_qnums) and len(got.Storage) > 3
    for x, y in zip(got.Storage, want.Storage):
        assert x.shape == y.shape and float((x - y).abs().max()) <= 1e-13 * float(y.abs().max())
    # the case the in-place test of ContractPlan must not fall for: A stored as [(k) ; (i)], so that A^T as second
    # operand needs exactly the stored bond order and row/column split -- but under exchanged tags
    u1 = [["U1", 0]]
    A2 = {"bonds": [{"dim": 3, "btype": 1, "qnums": [[1], [-1], [-2]], "syms": u1},
                    {"dim": 3, "btype": -1, "qnums": [[2], [1], [-2]], "syms": u1}], "labels": [20, 0], "rowrank": 1}
    B2 = {"bonds": [{"dim": 3, "btype": -1, "qnums": [[1], [-1], [-2]], "syms": u1},
                    {"dim": 2, "btype": 1, "qnums": [[1], [1]], "syms": u1}], "labels": [20, 40], "rowrank": 1}
    a2, b2 = psym(T, A2, seed=2, device=DEV), psym(T, B2, seed=501, device=DEV)
    c2 = T.Contract(a2, b2).Whole_transpose()
    ct2 = T.Contract(b2.Whole_transpose(), a2.Whole_transpose())
    c2.Permute([int(x) for x in ct2.labels], rowrank=int(ct2.rowrank), by_label=True)
    c2 = c2.Contiguous()
    ct2 = ct2 if ct2.is_contiguous() else ct2.Contiguous()
    assert np.array_equal(c2._block_qnums, ct2._block_qnums)
    for x, y in zip(ct2.Storage, c2.Storage):
        assert x.shape == y.shape and float((x - y).abs().max()) <= 1e-13 * max(1.0, float(y.abs().max()))
    # a transposed-back tensor equals the original bit for bit
    back = a.Whole_transpose().Whole_transpose()
    back.Permute([int(x) for x in a.labels], rowrank=int(a.rowrank), by_label=True)
    back = back.Contiguous()
    for x, y in zip(back.Storage, a.Storage):
        assert torch.equal(x, y)


def test_dense_blocks_and_shape_checks_follow_the_reference(T):
    """Found by tests/fuzz_vs_reference.py denseapi.  Dense GetBlock / PutBlock (UniTensor.py:3169-3195, :2941-2971):
    a rank-1 block when the row or the column space is empty, [rows, cols] otherwise, charges ignored.  Matmul with
    mismatching inner dimensions raises RuntimeError (torch.matmul in the reference) instead of running the GEMM with
    the wrong K; Svd labels its new bonds by the reference's position rule (linalg.py:535-539)."""
    dev = torch.device(DEV)
    x = T.UniTensor(bonds=[T.Bond(2), T.Bond(3), T.Bond(2)], rowrank=2, labels=[0, 1, 2], device=dev)
    x.Storage = torch.arange(12, dtype=torch.float64).reshape(2, 3, 2).to(dev)
    b = x.GetBlock()
    assert list(b.shape) == [6, 2] and b.rowrank == 1 and b.labels.tolist() == [0, 1]
    assert list(x.GetBlock(7).shape) == [6, 2]                      # charges are ignored on a dense tensor
    x.SetRowRank(0)
    b0 = x.GetBlock()
    assert list(b0.shape) == [12] and b0.rowrank == 0
    x.SetRowRank(3)
    b3 = x.GetBlock()
    assert list(b3.shape) == [12] and b3.rowrank == 1
    b3.Storage = b3.Storage * 2
    x.PutBlock(b3)
    assert torch.equal(x.Storage.cpu(), 2 * torch.arange(12, dtype=torch.float64).reshape(2, 3, 2))
    with pytest.raises(Exception):
        x.PutBlock(b)                                               # [6, 2] no longer is the block's shape
    m1 = T.UniTensor(bonds=[T.Bond(3), T.Bond(4)], rowrank=1, device=dev)
    m2 = T.UniTensor(bonds=[T.Bond(5), T.Bond(2)], rowrank=1, device=dev)
    with pytest.raises(RuntimeError):
        T.Matmul(m1, m2)
    with pytest.raises(RuntimeError):
        T.Chain_matmul(m1, m2)
    y = T.UniTensor(bonds=[T.Bond(3), T.Bond(4)], rowrank=1, labels=[5, -7], device=dev)
    y.Storage = torch.arange(12, dtype=torch.float64).reshape(3, 4).to(dev)
    u, s, v = T.Svd(y)
    assert u.labels.tolist() == [5, 0] and s.labels.tolist() == [0, -1] and v.labels.tolist() == [-1, -7]
