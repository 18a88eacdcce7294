"""CPU suite for the product's host side: the C-ABI library loads and exports every symbol the header
declares (no compute call is made without a GPU), Bond / Symmetry / Network / label algebra behave as
the reference's, the product refuses to compute without CUDA, and the N>1 sector gather works over gloo.
"""
import copy
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def T():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_t10_build", os.path.join(ROOT, "tor10_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()
    import tor10_b200
    return tor10_b200


def test_abi_exports_every_declared_symbol(T):
    hdr = open(os.path.join(ROOT, "include", "tor10_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(t10_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 13
    lib = ctypes.CDLL(T._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(T._lib.SIGNATURES), declared ^ set(T._lib.SIGNATURES)
    assert T._lib.load().t10_version() == 1
    assert ctypes.sizeof(T._lib.GemmProblem) == 48


def test_tile_planner_host_side(T):
    """t10_ggemm_plan_tiles is host code: heaviest tiles first, every output element covered once."""
    prob = np.zeros(3, dtype=T._lib.GEMM_PROBLEM_DTYPE)
    prob[0] = (0, 0, 0, 130, 70, 40, 40, 70, 70)
    prob[1] = (0, 0, 0, 5, 5, 500, 500, 5, 5)
    prob[2] = (0, 0, 0, 0, 9, 9, 9, 9, 9)
    for cfg, (bm, bn) in {0: (64, 64), 1: (128, 128), 2: (128, 64)}.items():
        n = ctypes.c_int64(0)
        lib = T._lib.load()
        assert lib.t10_ggemm_plan_tiles(cfg, 3, T._lib.np_ptr(prob), 0, 0, None, None, ctypes.byref(n)) == 0
        tiles = np.zeros((n.value, 4), dtype=np.int32)
        assert lib.t10_ggemm_plan_tiles(cfg, 3, T._lib.np_ptr(prob), 0, n.value, T._lib.np_ptr(tiles), None,
                                        ctypes.byref(n)) == 0
        want = -(-130 // bm) * -(-70 // bn) + 1
        assert n.value == want
        assert set(tiles[:, 0].tolist()) == {0, 1}  # empty problem produces no tile
        # LPT slot table: every tile exactly once, slot-major, slots cover [0, ntiles)
        for nslots in (1, 2, 5):
            t2 = np.zeros((n.value, 4), dtype=np.int32)
            ss = np.zeros(nslots + 1, dtype=np.int32)
            assert lib.t10_ggemm_plan_tiles(cfg, 3, T._lib.np_ptr(prob), nslots, n.value, T._lib.np_ptr(t2),
                                            T._lib.np_ptr(ss), ctypes.byref(n)) == 0
            assert ss[0] == 0 and ss[-1] == n.value and (np.diff(ss) >= 0).all()
            assert sorted(map(tuple, t2.tolist())) == sorted(map(tuple, tiles.tolist()))
        b0, b1 = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.t10_ggemm_tile_shape(cfg, ctypes.byref(b0), ctypes.byref(b1)) == 0 and (b0.value, b1.value) == (bm, bn)


def _streamk_tables(T, prob, nslots, phases, monkeypatch):
    """GemmGroup.build_streamk is host code (numpy + the host-side tile planner): run it without a device."""
    monkeypatch.setenv("TOR10_GGEMM_STREAMK", "force")
    g = T._plan.GemmGroup.__new__(T._plan.GemmGroup)
    g.problems, g.nprob = prob, len(prob)
    g.d_prob = torch.zeros(48, dtype=torch.uint8)
    assert g.build_streamk(nslots, phases=phases)
    return g, g.d_sk_tiles.numpy(), g.d_sk_pieces.numpy(), g.d_sk_slots.numpy()


@pytest.mark.parametrize("phases", [1, 2, 3])
@pytest.mark.parametrize("nslots", [7, 40, 296])
def test_streamk_piece_tables(T, nslots, phases, monkeypatch):
    """The stream-K schedule (one or several phases): every tile's k-steps are covered exactly once, a parked
    piece owns its park slot, the finishing piece lists exactly its tile's parked slots in k order, and running the
    CTAs' lists in order never deadlocks (a waiting piece only waits for pieces that can run before it)."""
    rng = np.random.default_rng(nslots * 10 + phases)
    prob = np.zeros(9, dtype=T._lib.GEMM_PROBLEM_DTYPE)
    for i in range(len(prob)):
        m, n, k = int(rng.integers(20, 400)), int(rng.integers(20, 400)), int(rng.integers(16, 1500))
        prob[i] = (0, 0, 0, m, n, k, k + (k & 1), n + (n & 1), n + (n & 1))
    g, tiles, pcs, starts = _streamk_tables(T, prob, nslots, phases, monkeypatch)
    assert starts[0] == 0 and starts[-1] == len(tiles) == len(pcs) and len(starts) == nslots + 1
    kts = (prob["k"][tiles[:, 0]].astype(np.int64) + 15) // 16
    by_tile = {}
    for t in range(len(tiles)):
        by_tile.setdefault(tuple(tiles[t]), []).append((int(pcs[t, 0]), int(pcs[t, 1]), int(pcs[t, 2]), int(pcs[t, 3]), int(kts[t])))
    # the planner's own tile list, each tile exactly once
    n = ctypes.c_int64(0)
    lib = T._lib.load()
    assert lib.t10_ggemm_plan_tiles(46, len(prob), T._lib.np_ptr(prob), 0, 0, None, None, ctypes.byref(n)) == 0
    assert len(by_tile) == n.value
    slots_seen = set()
    for key, ps in by_tile.items():
        ps.sort()
        assert ps[0][0] == 0 and ps[-1][1] == ps[-1][4]
        assert all(a[1] == b[0] for a, b in zip(ps, ps[1:])), "k ranges must tile [0, KT) without gaps or overlap"
        assert all(a[1] > a[0] for a in ps)
        parked = [a[2] for a in ps[:-1]]
        assert all(x >= 0 and x != nslots for x in parked) and ps[-1][2] == -1
        assert not (set(parked) & slots_seen)
        slots_seen |= set(parked)
        nfix, first = ps[-1][3] & 0xff, ps[-1][3] >> 8
        assert nfix == len(parked) and parked == list(range(first, first + nfix))
    assert g.sk_parked == len(slots_seen) and max(slots_seen, default=0) < len(g.d_sk_flags)
    assert (max(slots_seen, default=0) + 1) * 128 * 34 <= g.d_sk_ws.numel()
    # execution: CTAs advance through their lists; a finishing piece needs its parked slots raised
    pc = starts[:-1].astype(np.int64).copy()
    raised, progress = set(), True
    while progress:
        progress = False
        for c in range(nslots):
            while pc[c] < starts[c + 1]:
                a, b, slot, w = (int(x) for x in pcs[pc[c]])
                need = set(range(w >> 8, (w >> 8) + (w & 0xff)))
                if slot < 0 and not need <= raised:
                    break
                if slot >= 0:
                    raised.add(slot)
                pc[c] += 1
                progress = True
    assert (pc == starts[1:]).all(), "the schedule deadlocks"
    # equal shares: no CTA holds much more than the mean
    load = np.array([sum(int(pcs[t, 1] - pcs[t, 0]) + 2 for t in range(starts[c], starts[c + 1])) for c in range(nslots)])
    if load.mean() >= 16 * g.sk_phases:
        assert load.max() <= 1.25 * load.mean() + 4 * g.sk_phases


def test_tile_cfg_selection_rules(T, monkeypatch):
    """_plan.pick_tile_cfg (the kernel chosen for an fp64 group) as a decision table; SM count pinned to 148."""
    monkeypatch.setattr(T._plan, "_sm_count", lambda: 148)
    for k in ("TOR10_GGEMM_CFG", "TOR10_SHARDED_SK"):
        monkeypatch.delenv(k, raising=False)

    def group(shapes, odd=False):
        pr = np.zeros(len(shapes), dtype=T._lib.GEMM_PROBLEM_DTYPE)
        for i, (m, n, k) in enumerate(shapes):
            pr[i] = (0, 0, 0, m, n, k, k + (1 if odd else 0), n, n)
        return pr

    pick = T._plan.pick_tile_cfg
    c3a_like = group([(400, 400, 400)] * 40)
    assert pick(None) == 3 and pick(group([])) == 3
    assert pick(c3a_like, base_aligned=True) == 46                     # block-sparse group: mixed shapes (+ stream-K)
    assert pick(c3a_like, base_aligned=False) == 3                     # operand base not 16-byte aligned
    assert pick(group([(400, 400, 400)] * 40, odd=True), base_aligned=True) == 3   # odd leading dimension
    assert pick(group([(2048, 2048, 2048)]), base_aligned=True) == 40  # one dense GEMM, few waves
    assert pick(group([(8192, 8192, 64)] * 4), base_aligned=True) == 42   # many waves of 64x128 tiles
    # sharded ranks: stream-K while every CTA keeps 16 k-steps, else 64x64, else 32x32 for a handful of tiles
    assert pick(c3a_like, sharded=True, base_aligned=True) == 46
    assert pick(group([(400, 400, 32)] * 40), sharded=True, base_aligned=True) == 40
    assert pick(group([(100, 100, 64)] * 10), sharded=True, base_aligned=True) == 8
    assert pick(group([(100, 100, 64)] * 10), sharded=True, base_aligned=False) == 8
    monkeypatch.setenv("TOR10_SHARDED_SK", "0")
    assert pick(c3a_like, sharded=True, base_aligned=True) == 40
    monkeypatch.setenv("TOR10_GGEMM_CFG", "2")
    assert pick(c3a_like, base_aligned=True) == 2


def test_no_cpu_compute_path(T):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    b1 = T.Bond(2, T.BD_KET, qnums=[[0], [1]])
    b2 = T.Bond(2, T.BD_BRA, qnums=[[0], [1]])
    with pytest.raises(RuntimeError):
        T.UniTensor(bonds=[b1, b2], rowrank=1)                   # block map is a CUDA kernel
    x = T.UniTensor(bonds=[T.Bond(2), T.Bond(3)], rowrank=1)
    y = T.UniTensor(bonds=[T.Bond(3), T.Bond(2)], rowrank=1, labels=[1, 2])
    with pytest.raises(RuntimeError):
        T.Contract(x, y)
    with pytest.raises(RuntimeError):
        T.Matmul(x, y)


def test_product_never_imports_the_oracle():
    """oracle/ and tests/abi_emulator.py are test infrastructure: no file of the package (or the example script)
    may import or mention them."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("tor10_b200", "examples"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "abi_emulator" in txt \
                            or "emu_plugin" in txt:
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_bond_matches_reference_goldens(T, golden):
    z, specs = golden("bond")
    for key, sp in specs.items():
        syms = [T.Symmetry.U1() if s[0] == "U1" else T.Symmetry.Zn(s[1]) for s in sp["syms"]]
        a = T.Bond(len(sp["raw_a"]), T.BD_KET, qnums=sp["raw_a"], sym_types=syms)
        b = T.Bond(len(sp["raw_b"]), T.BD_KET, qnums=sp["raw_b"], sym_types=syms)
        assert np.array_equal(a.qnums, z[key + "/a_sorted"]) and np.array_equal(b.qnums, z[key + "/b_sorted"])
        assert np.array_equal(a.GetUniqueQnums(), z[key + "/a_unique"])
        assert np.array_equal(T._fx_GetCommRows(a.GetUniqueQnums(), b.GetUniqueQnums()), z[key + "/comm"])
        c = copy.deepcopy(a)
        c.combine(b)
        assert c.dim == a.dim * b.dim and np.array_equal(c.qnums, z[key + "/combined"])
        assert a.qnums is copy.deepcopy(a).qnums             # copies share the immutable table
        assert a.signature() == copy.deepcopy(a).signature()
        assert (c.signature() == a.signature()) == (c.dim == a.dim and np.array_equal(c.qnums, a.qnums))


def test_bond_api_errors_and_docstring_kats(T):
    with pytest.raises(Exception):
        T.Bond(0)
    with pytest.raises(Exception):
        T.Bond(2, T.BD_REG, qnums=[[0], [1]])                 # Bond.py:240-241
    with pytest.raises(ValueError):
        T.Bond(3, T.BD_KET, qnums=[[0], [1]])
    with pytest.raises(TypeError):
        T.Bond(2, T.BD_KET, qnums=[[0], [5]], sym_types=[T.Symmetry.Zn(3)])
    with pytest.raises(ValueError):
        T.Symmetry.Zn(1)
    a = T.Bond(3, T.BD_BRA, qnums=[[0], [0], [1]])
    b = T.Bond(4, T.BD_BRA, qnums=[[0], [2], [-3], [-1]])
    assert a.qnums.flatten().tolist() == [1, 0, 0] and b.qnums.flatten().tolist() == [2, 0, -1, -3]
    a.combine(b)                                               # Bond.py:341-347
    assert a.qnums.flatten().tolist() == [3, 1, 0, -2, 2, 0, -1, -3, 2, 0, -1, -3]
    z1 = T.Bond(3, T.BD_BRA, qnums=[[0], [2], [1]], sym_types=[T.Symmetry.Zn(4)])
    z1 * -1                                                    # in-place flip, no mod (Bond.py:469-477)
    assert z1.qnums.flatten().tolist() == [-2, -1, 0]
    assert z1.GetDegeneracy(-1) == 1
    assert (T.Bond(3) == T.Bond(3)) and not (T.Bond(3) == T.Bond(4))
    with pytest.raises(ValueError):
        _ = T.Bond(3) == 3
    with pytest.raises(ValueError):
        a.qnums[0, 0] = 7                                      # immutable: keeps plan-cache keys honest


def test_dense_metadata_matches_reference_tests(T):
    """tests/test_UniTensor.py:10-81 of the reference (labels, Reshape, Permute) -- metadata only."""
    cpu = torch.device("cpu")
    os.environ["TOR10_B200_FORCE_CUDA"] = "0"
    try:
        ut = T.UniTensor(bonds=[T.Bond(3), T.Bond(4), T.Bond(5)], rowrank=2, labels=[10, 11, 12], device=cpu)
        ut.SetLabel(2, 1)
        assert ut.labels.tolist() == [10, 2, 12]
        ut.SetLabels([3, 4, 5])
        assert ut.labels.tolist() == [3, 4, 5]
        with pytest.raises(ValueError):
            ut.SetLabels([1, 1, 2])
        r = ut.Reshape([12, 5], rowrank=1)
        assert tuple(r.shape) == (12, 5) and r.rowrank == 1 and r.labels.tolist() == [0, 1]
        ut.Permute([2, 0, 1], rowrank=1)
        assert tuple(ut.shape) == (5, 3, 4) and ut.labels.tolist() == [5, 3, 4] and ut.rowrank == 1
        assert not ut.is_contiguous()
        ut.Permute([3, 4, 5], by_label=True)
        assert tuple(ut.shape) == (3, 4, 5)
        with pytest.raises(Exception):
            T.UniTensor(bonds=[T.Bond(3), T.Bond(4)], labels=[1, 1], rowrank=1, device=cpu)
        with pytest.raises(Exception):
            T.UniTensor(bonds=[T.Bond(3), T.Bond(4)], device=cpu)          # rowrank required for untagged
        d = T.UniTensor(bonds=[T.Bond(3), T.Bond(3)], rowrank=1, is_diag=True, device=cpu)
        assert tuple(d.shape) == (3, 3) and d.Storage.shape == (3,)
        tg = T.UniTensor(bonds=[T.Bond(2, T.BD_KET), T.Bond(2, T.BD_BRA)], device=cpu)
        assert tg.rowrank == 1 and tg.is_braket_form()
    finally:
        os.environ.pop("TOR10_B200_FORCE_CUDA", None)


def test_network_parser_and_checks(T, golden):
    z, specs = golden("dense")
    N = T.Network()
    N.Fromstring(specs["net0"]["text"])
    assert list(N.tensors) == ["A", "B", "C"] and N.Order == "(A,B),C"
    assert N.TOUT[0].tolist() == [-1, -2] and N.TOUT[1].tolist() == [3, 4]
    with pytest.raises(Exception):
        N.Launch()                                             # nothing put (Network.py:397-398)
    with pytest.raises(ValueError):
        N.Put("Q", T.UniTensor(bonds=[T.Bond(2), T.Bond(2)], rowrank=1))
    with pytest.raises(TypeError):
        N.Put("B", T.UniTensor(bonds=[T.Bond(2), T.Bond(2), T.Bond(2)], rowrank=1))
    for bad in ("A: 1 2; 3\n", "A 1;2\nTOUT: 1;2\n", "A: 1;2\nA: 2;3\nTOUT: 1;3\n", "A: 1;2\nTOUT: ;\nOrder: ((A)\n"):
        with pytest.raises(Exception):
            T.Network().Fromstring(bad)


def test_label_match_is_label_sorted(T):
    same, ai, bi, af, bf = T._plan.label_match(np.array([5, 30, 7, 21]), np.array([21, 9, 30]))
    assert same.tolist() == [21, 30] and ai.tolist() == [3, 1] and bi.tolist() == [0, 2]   # quirk Q5
    assert af.tolist() == [0, 2] and bf.tolist() == [1]


def test_partition_contiguous(T):
    from tor10_b200.parallel import partition_contiguous
    cost = np.array([1, 1, 8, 1, 1, 1, 9, 1, 1], dtype=float)
    b = partition_contiguous(cost, 3)
    assert b[0] == 0 and b[-1] == len(cost) and len(b) == 4 and all(x <= y for x, y in zip(b, b[1:]))
    assert max(cost[b[i]:b[i + 1]].sum() for i in range(3)) <= 11.0
    assert partition_contiguous(np.ones(2), 4) == [0, 1, 2, 2, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tor10_b200 import parallel
    parallel.enable()
    assert parallel.current_shard() == (rank, world)

    class L:   # the three layout fields the gather needs
        arena_off = np.array([0, 16, 48, 64])
        arena_elems, nsec = 96, 4

    bounds = parallel.partition_contiguous(np.array([3.0, 1.0, 1.0, 3.0]), world)
    ranges = parallel.arena_ranges(L, bounds)
    arena = torch.zeros(96 + 16, dtype=torch.float64)
    lo, hi = ranges[rank]
    arena[lo:hi] = rank + 1.0                      # "my sectors"
    parallel.gather_ranges(arena, ranges, rank)
    want = torch.zeros_like(arena)
    for r, (lo, hi) in enumerate(ranges):
        want[lo:hi] = r + 1.0
    q.put((rank, bool(torch.equal(arena, want)), bounds))
    dist.destroy_process_group()


def test_sector_gather_world2_gloo(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2] == [0, 2, 4]


# ---- round 2: bounded plan caches, sector-resident ownership logic (pure host code) ----------------------------
def test_plan_cache_is_a_bounded_lru(T):
    from tor10_b200 import _plan

    class P:
        def __init__(self, nbytes):
            self.nbytes = nbytes

        def plan_bytes(self):
            return self.nbytes
    c = _plan.LRU(3)
    for k in range(5):
        c[k] = P(10)
    assert len(c) == 3 and c.get(0) is None and c.get(1) is None and c.get(4) is not None and c.evicted == 2
    c.get(2)                       # touched: 3 is now the oldest
    c[5] = P(10)
    assert c.get(3) is None and c.get(2) is not None
    b = _plan.LRU(100, max_bytes=25)
    b["a"], b["b"], b["c"] = P(10), P(10), P(10)       # 30 bytes: the oldest goes
    assert b.get("a") is None and b.bytes == 20
    b["big"] = P(1000)                                  # a single entry over the budget is kept (it is in use)
    assert len(b) == 1 and b.get("big") is not None
    # GEMM groups and contraction plans report the device tables they hold (stream-K workspace, tensor maps)

    class G:
        def __init__(self):
            self.ws = torch.zeros(10, dtype=torch.float64)
            self.maps = {1: torch.zeros(256, dtype=torch.uint8), 2: "x"}
            self.keep = [torch.zeros(3, dtype=torch.int32), None]
            self.n = 5
    assert _plan._owned_tensor_bytes(G()) == 80 + 256 + 12
    import sys
    U = sys.modules["tor10_b200.UniTensor"]
    assert isinstance(U._GEMM_GROUPS, _plan.LRU) and hasattr(_plan.GemmGroup, "plan_bytes") \
        and hasattr(_plan.ContractPlan, "plan_bytes")


def test_pinned_objects_survive_eviction(T):
    """graph.StaticLoop keeps what _plan.pinned_objects() returned: a later LRU eviction or clear_caches() must not
    free a plan a captured graph still addresses by raw pointer."""
    import gc
    import weakref
    from tor10_b200 import _plan

    class P:
        def plan_bytes(self):
            return 8
    saved = dict(_plan._CONTRACTS.d)
    try:
        p = P()
        ref = weakref.ref(p)
        _plan._CONTRACTS[("pin-test",)] = p
        del p
        pins = _plan.pinned_objects()
        assert any(o is ref() for o in pins)
        _plan.clear_caches()
        gc.collect()
        assert ref() is not None                     # held by the pins
        del pins
        gc.collect()
        assert ref() is None
    finally:
        for k, v in saved.items():
            _plan._CONTRACTS[k] = v


def test_resident_ownership_and_fetch_plan(T):
    """Ownership propagates from a resident first operand; a rank pulls exactly the needed sectors it does
    not hold, grouped by owner, as whole 16-byte vectors."""
    from tor10_b200 import parallel
    a_owner = np.array([0, 1, -1, 1, 0])
    matched = [(0, 4, 0), (1, 0, 1), (3, 1, 2), (4, 2, 3)]            # (out sector, A sector, B sector)
    own = parallel.inherit_owner(a_owner, matched, 6, None, 2)
    assert own.tolist() == [0, 0, -1, 1, -1, -1]                     # out 4 pairs with a zero A sector: nobody's
    res0 = parallel.Residency(2, 0, own, 7, None)
    assert res0.valid.tolist() == [True, True, True, False, True, True]

    class L:
        arena_off = np.array([0, 16, 48, 64, 96, 112])
        rows = np.array([2, 4, 1, 3, 1, 1])
        ld = np.array([4, 6, 2, 6, 2, 2])
    plan = parallel.fetch_plan(L, res0, [0, 3, 5])
    assert plan == {1: [(64, 84)]}                                    # 3 * 6 = 18 elements -> 20 (whole vectors)
    res1 = parallel.Residency(2, 1, own, 7, None)
    assert parallel.fetch_plan(L, res1, [0, 1, 3]) == {0: [(0, 8), (16, 40)]}
    assert res0.key() == res1.key()                                   # the plan key does not depend on the rank


def _collective_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tor10_b200 import parallel
    parallel.enable()

    class L:
        arena_off = np.array([0, 16, 48, 64])
        rows = np.array([2, 4, 2, 4])
        ld = np.array([8, 8, 8, 8])
        nsec = 4

    class Tn:      # the three fields gather_collective touches
        pass
    t = Tn()
    t._layout = L
    owner = np.array([0, 0, 1, -1])
    t._res = parallel.Residency(world, rank, owner, 0, None)
    t._arena = torch.zeros(96 + 16, dtype=torch.float64)
    for s_ in range(4):
        if owner[s_] == rank:
            lo = int(L.arena_off[s_])
            t._arena[lo:lo + int(L.rows[s_] * L.ld[s_])] = 10.0 * (rank + 1) + s_
    parallel.gather_collective(t)
    want = torch.zeros_like(t._arena)
    for s_ in range(4):
        if owner[s_] >= 0:
            lo = int(L.arena_off[s_])
            want[lo:lo + int(L.rows[s_] * L.ld[s_])] = 10.0 * (owner[s_] + 1) + s_
    q.put((rank, bool(torch.equal(t._arena, want)) and bool(t._res.valid.all())))
    dist.destroy_process_group()


def test_resident_collective_completion_world2_gloo(T):
    """A resident tensor outside peer-readable memory is completed by one broadcast per owner run (the N > 1 path of
    the sharded Svd_truncate's factors), here over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_collective_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_bench_reads_roofline_traffic_from_the_committed_capture(T, tmp_path, monkeypatch):
    """bench.py takes `roofline.traffic` from profiles/r02_ncu_raw_final.csv and refuses a capture that does not hold
    the kernel the roofline names."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    got = bench.ncu_traffic("k_ggemm_f64_tma")
    assert got is not None and 4.0e7 < got < 9.0e7            # operands once + parked partials, C stays in L2
    rel = bench.ncu_traffic("k_relayout_segments")
    assert rel is not None and 2.0e7 < rel < 4.0e7            # the source once: no index traffic
    with pytest.raises(AssertionError):
        bench.ncu_traffic("k_some_other_kernel")
    monkeypatch.setattr(bench, "NCU_RAW", str(tmp_path / "missing.csv"))
    assert bench.ncu_traffic("k_ggemm_f64_tma") is None


def test_dense_shard_threshold(T, monkeypatch):
    """Dense GEMMs are split over the ranks only above the flop threshold and when every rank gets a 64-row tile."""
    import sys
    U = sys.modules["tor10_b200.UniTensor"]
    from tor10_b200 import parallel
    x = torch.zeros(1, dtype=torch.float64)
    assert U._dense_shard(2048, 2048, 2048, x) is None                     # sharding is off
    monkeypatch.setitem(parallel._STATE, "enabled", True)
    monkeypatch.setitem(parallel._STATE, "shard", (1, 4))
    assert U._dense_shard(2048, 2048, 2048, x) == (1, 4)
    assert U._dense_shard(128, 2048, 2048, x) is None                      # fewer than one tile of rows per rank
    assert U._dense_shard(256, 256, 256, x) is None                        # below TOR10_DENSE_SHARD_MIN_GFLOP
    assert U._dense_shard(2048, 2048, 2048, x.float()) is None             # fp64 only


def test_pull_sectors_host_logic(T, monkeypatch):
    """parallel.pull_sectors on host memory with a stand-in for the gather kernel: the right slices of the right
    owners arrive, neighbouring sectors of one owner merge into one slice, the argument arrays are cached on the slot
    (a list request and a range request with the same two numbers do not share an entry), validity and the slot's
    write-after-read bound are updated.  (The kernel itself: tests/test_gpu_multi.py.)"""
    import ctypes as C
    from tor10_b200 import parallel, _lib

    class L:
        arena_off = np.array([0, 16, 48, 64, 96])
        rows = np.array([2, 4, 2, 4, 2])
        ld = np.array([8, 8, 8, 8, 8])
    n = 112
    peers = [np.full(n, 100.0 * (r + 1)) + np.arange(n) for r in range(3)]          # rank r's copy of the arena
    mine = torch.zeros(n, dtype=torch.float64)

    class Pool:
        bufs = [mine]
        peer_ptrs = [[p.ctypes.data for p in peers]]
    slot = parallel._Slot(Pool(), 0)
    owner = np.array([0, 1, 1, 2, -1])
    res = parallel.Residency(3, 0, owner, 5, slot)

    class Tns:
        _res, _layout, _arena = res, L, mine
    calls = []

    class Lib:
        def t10_gather_peers(self, dt, dst, nseg, ptrs, los, his, st):
            out = np.ctypeslib.as_array((C.c_double * n).from_address(dst))
            for i in range(nseg):
                src = np.ctypeslib.as_array((C.c_double * n).from_address(ptrs[i]))
                out[los[i]:his[i]] = src[los[i]:his[i]]
            calls.append([(ptrs[i], los[i], his[i]) for i in range(nseg)])
            return 0

    class RW:
        world, rank, seq = 3, 0, 9
        waited = []

        def wait(self, need):
            self.waited.append(list(need))
    rw = RW()
    monkeypatch.setattr(parallel, "resident_world", lambda device=None: rw)
    monkeypatch.setattr(_lib, "load", lambda: Lib())
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    parallel.pull_sectors(Tns, [0, 1, 2, 4])             # 0 and 4 are valid here already; 1 and 2 live on rank 1
    assert rw.waited == [[0, 5, 0]]
    assert calls == [[(peers[1].ctypes.data, 16, 64)]]                       # sectors 1+2: one slice
    assert np.array_equal(mine.numpy()[16:64], peers[1][16:64]) and mine.numpy()[64:].sum() == 0
    assert res.valid.tolist() == [True, True, True, False, True] and slot.last_read == 10
    assert len(slot.pull_args) == 1
    parallel.pull_sectors(Tns, [0, 1, 2, 4])             # nothing left to pull: no wait, no launch
    assert len(calls) == 1 and len(rw.waited) == 1
    parallel.pull_sectors(Tns, range(3, 4))              # sector 3 from rank 2
    assert calls[-1] == [(peers[2].ctypes.data, 64, 96)] and res.valid.all()
    assert np.array_equal(mine.numpy()[64:96], peers[2][64:96])
    assert set(k[0] for k in slot.pull_args) == {(0, 1, 2, 4), ("range", 3, 4, 1)}
    Tns._res = parallel.Residency(3, 0, owner, 6, None)                     # not in peer-readable memory any more
    with pytest.raises(RuntimeError):
        parallel.pull_sectors(Tns, [3])


def test_materialize_and_ring_reuse_order(T, monkeypatch):
    """parallel.materialize: the private copy is taken BEFORE the read is published (afterwards a pushing peer may
    overwrite the slot), the slot's write-after-read bound moves to the published value and the slot is released;
    ArenaPool.acquire gathers a tensor still living in the slot it hands out."""
    import weakref
    from tor10_b200 import parallel
    log = []

    class Arena:
        def clone(self):
            log.append("clone")
            return self

    class RW:
        world, rank, seq = 2, 0, 4

        def next_seq(self):
            self.seq += 1
            return self.seq

        def signal(self, v):
            log.append(("signal", v))

        def wait(self, need):
            log.append(("wait", list(need)))
    rw = RW()
    monkeypatch.setattr(parallel, "resident_world", lambda device=None: rw)

    class Pool:
        depth = 2
    pool = Pool()
    pool.slots = [parallel._Slot(pool, 0), parallel._Slot(pool, 1)]
    pool.turn = 0
    owner = np.array([0, -1])

    class Tns:
        def __init__(self, slot):
            self._res = parallel.Residency(2, 0, owner, 3, slot)
            self._arena, self._blocks, self._layout = Arena(), "blocks", None

        def _materialize(self):
            parallel.materialize(self)
    t = Tns(pool.slots[0])
    pool.slots[0].holder = weakref.ref(t)
    parallel.materialize(t)
    assert log == ["clone", ("signal", 5)]                  # every sector valid here: no pull, no wait
    assert t._res is None and t._blocks is None and pool.slots[0].holder is None and pool.slots[0].last_read == 5
    parallel.materialize(t)                                  # idempotent
    assert len(log) == 2
    # ring reuse: slot 0 is free, slot 1 still holds a live tensor -> it is gathered out when the ring comes round
    u = Tns(pool.slots[1])
    pool.slots[1].holder = weakref.ref(u)
    s0, war0 = parallel.ArenaPool.acquire(pool)
    assert s0 is pool.slots[0] and war0 == 5 and len(log) == 2
    s1, war1 = parallel.ArenaPool.acquire(pool)
    assert s1 is pool.slots[1] and u._res is None and log[2:] == ["clone", ("signal", 6)] and s1.holder is None
    assert pool.turn == 0


def test_resident_wait_is_skipped_when_already_acquired(T, monkeypatch):
    """ResidentWorld.wait launches the acquire kernel only for progress values the current stream has not acquired
    yet: nothing to wait for, this rank's own (stream-ordered) operations and repeated waits cost no launch."""
    from tor10_b200 import parallel, _lib
    rw = object.__new__(parallel.ResidentWorld)
    rw.world, rw.rank, rw.seq, rw.acquired = 3, 1, 7, {}
    rw.word_ptrs = None

    class Err:
        def data_ptr(self):
            return 0
    rw.err = Err()
    calls = []

    class Lib:
        def t10_wait_flags(self, n, ptrs, need, timeout_ms, err, st):
            calls.append(([int(need[i]) for i in range(n)], timeout_ms, st))
            return 0
    stream = [11]
    monkeypatch.setattr(_lib, "load", lambda: Lib())
    monkeypatch.setattr(_lib, "stream_ptr", lambda: stream[0])
    assert rw.wait([0, 0, 0]) is False and not calls
    assert rw.wait([0, 7, 0]) is False and not calls            # own operations up to seq: stream order
    assert rw.wait([5, 0, 0]) is True and calls[-1][0] == [5, 0, 0] and calls[-1][1] == parallel.WAIT_TIMEOUT_MS
    assert rw.wait([5, 3, 0]) is False and len(calls) == 1      # known on this stream
    assert rw.wait([4, 0, 6]) is True and calls[-1][0] == [4, 0, 6]
    assert rw.acquired[11] == [5, 7, 6]
    stream[0] = 12                                              # another stream has acquired nothing yet
    assert rw.wait([5, 0, 0]) is True and len(calls) == 3 and calls[-1][2] == 12


def test_resident_segments_pick_the_owner_arena(T):
    """_plan._resident_segments: the `source` column of a relayout's segment table is 0 for pieces whose source
    sector is valid here and k > 0 for the k-th foreign owner (read through its NVLink mapping inside the copy);
    zero-fill pieces (src < 0) stay local; cached per (owner, valid) state."""
    from tor10_b200 import _plan, parallel

    class M:
        arena_off = np.array([0, 100, 200, 300])

    class Rel:
        pass
    rel = Rel()
    rel.M = M
    #                        dst  src  len source
    rel.d_seg = torch.tensor([[0, 10, 8, 0], [8, 150, 8, 0], [16, 250, 4, 0], [20, 305, 6, 0], [26, -1, 2, 0],
                              [28, 199, 1, 0]], dtype=torch.int32)

    class Plan:
        _res_segs = {}
    owner = np.array([0, 2, 1, 2])
    res = parallel.Residency(3, 0, owner, 1, None)            # rank 0 holds sector 0 only

    class Tns:
        _res = res
    seg, foreign = _plan._resident_segments(Plan, "a", rel, Tns)
    assert foreign == [1, 2]
    assert seg[:, 3].tolist() == [0, 2, 1, 2, 0, 2]           # sector 1 and 3 from rank 2 (2nd foreign), 2 from rank 1
    assert seg[:, :3].tolist() == rel.d_seg[:, :3].tolist() and rel.d_seg[:, 3].sum() == 0    # the plan's table is untouched
    assert _plan._resident_segments(Plan, "a", rel, Tns)[0] is seg                             # cached
    res.valid[1] = True                                        # sector 1 has been pulled: read locally from now on
    seg2, foreign2 = _plan._resident_segments(Plan, "a", rel, Tns)
    assert seg2[:, 3].tolist() == [0, 0, 1, 2, 0, 0] and foreign2 == [1, 2]
    rel.d_seg = None
    assert _plan._resident_segments(Plan, "a", rel, Tns) is None
