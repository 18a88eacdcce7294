"""linalg (API of tor10/linalg.py).

`Matmul` / `Chain_matmul` run on the family-3 GEMM kernel (reference: torch.matmul / chain_matmul,
linalg.py:722-733, :806).  The decompositions stay on torch / cuSOLVER on the device, as the north
star prescribes (`Svd_truncate` is timed separately by bench.py); they keep the reference's label
and bond conventions.
"""
import copy
import os

import numpy as np
import torch

from . import _lib
from .Bond import Bond
from .UniTensor import Contract, UniTensor, _dense_gemm


def _new_label(a):
    """Smallest negative label of `a` (0 if none): the new internal bonds of the symmetric SVD get tmp-1, tmp-2."""
    neg = a.labels[a.labels < 0]
    return 0 if len(neg) == 0 else int(np.min(neg))


def _new_label_ref(a):
    """The reference's rule for the dense factorisations (linalg.py:363-367, :428-432, :535-539): the POSITION of the
    first negative label (`np.min(np.argwhere(labels < 0))`, 0 if there is none), so the new bonds are labelled
    -1 / -2 or 0 / -1 whatever the labels' values.  Kept as is: scripts relabel the factors anyway (iTEBD.py:118)."""
    neg = np.argwhere(a.labels < 0)
    return 0 if len(neg) == 0 else int(np.min(neg))


def _rank2(x, bonds=None, labels=None, is_diag=False):
    if bonds is None:
        n = x.shape[0]
        bonds = [Bond(n), Bond(n)] if is_diag else [Bond(x.shape[0]), Bond(x.shape[1])]
    t = UniTensor(bonds=bonds, rowrank=1, labels=labels, check=False, is_diag=is_diag)
    t._mac(torch_tensor=x)
    return t


def _scalar(x):
    t = UniTensor(bonds=[], labels=[], rowrank=0, check=False)
    t._mac(torch_tensor=x)
    return t


def Hosvd(a, order, bonds_group, by_label=False, core=True):
    """linalg.py:6-138 (dense tensors): factors U_g of each bond group and, optionally, the core."""
    if not isinstance(a, UniTensor):
        raise TypeError("Hosvd(UniTensor,*args)", "[ERROR] the input should be a UniTensor")
    if not (isinstance(bonds_group, list) or isinstance(bonds_group, np.ndarray)):
        raise TypeError("Hosvd(UniTensor,order,bonds_group,*args)",
                        "[ERROR] the bonds_group should be a python list or 1d numpy array")
    if not (isinstance(order, list) or isinstance(order, np.ndarray)):
        raise TypeError("Hosvd(UniTensor,order,bonds_group,*args)",
                        "[ERROR] the order should be a python list or 1d numpy array")
    if len(order) != len(a.labels):
        raise ValueError("Hosvd", "[ERROR] the size of order should be equal to the rank of input UniTensor.")
    if len(a.labels) < 3:
        raise Exception("Hosvd", "[ERROR], Hosvd can only perform on a UniTensor with rank > 2. For a rank-2 tensor, "
                                 "using Svd() instead.")
    if any(int(x) <= 0 for x in bonds_group):
        raise ValueError("Hosvd", "[ERROR] bonds_group cannot have elements <=0")
    if a.is_symm:
        raise Exception("Hosvd can only operate on non-symmetry tensor")
    if sum(int(x) for x in bonds_group) != len(a.labels):
        raise ValueError("Hosvd", "[ERROR] bonds_group must cover all the bonds")
    lbls = np.array(order, dtype=np.int64) if by_label else a.labels[np.array(order, dtype=np.int64)]
    work = copy.deepcopy(a)
    work.Permute(lbls, by_label=True)
    start = int(np.min(a.labels))
    start = start - 1 if start <= 0 else -1
    factors, pos = [], 0
    x = UniTensor._contiguous_dense(work._dense)
    for g in bonds_group:
        g = int(g)
        front = list(range(pos, pos + g))
        rest = [i for i in range(x.dim()) if i not in front]
        mat = UniTensor._contiguous_dense(x.permute(front + rest)).reshape(
            int(np.prod([x.shape[i] for i in front])), -1)
        u, _, _ = torch.linalg.svd(mat, full_matrices=False)
        f = UniTensor(bonds=[Bond(int(x.shape[i])) for i in front] + [Bond(u.shape[-1])],
                      labels=list(lbls[pos:pos + g]) + [start], rowrank=g, check=False)
        f._mac(torch_tensor=u.reshape([int(x.shape[i]) for i in front] + [u.shape[-1]]))
        factors.append(f)
        start -= 1
        pos += g
    if not core:
        return factors
    out = work
    for f in factors:
        out = Contract(out, f)
    return factors, out


def Abs(a):
    if not isinstance(a, UniTensor):
        raise TypeError("Abs(UniTensor)", "[ERROR] the input should be a UniTensor")
    if a.is_symm:
        return a._like(torch.abs(a._arena_tensor()))
    return a._like(torch.abs(a._dense))


def Mean(a):
    if not isinstance(a, UniTensor):
        raise TypeError("Mean(UniTensor)", "[ERROR] the input should be a UniTensor")
    if a.is_symm:
        raise Exception("Mean(UniTensor)", "[ERROR] cannot get mean for a symmetry tensor. GetBlock first")
    return _scalar(torch.mean(a._dense))


def Otimes(a, b):
    """Kronecker product of two rank-2 untagged tensors (linalg.py:200-256)."""
    if not (isinstance(a, UniTensor) and isinstance(b, UniTensor)):
        raise TypeError("Otimes", "[ERROR], Otimes only accept UniTensor as arguments.")
    if not (len(a.labels) == 2 and len(b.labels) == 2):
        raise TypeError("Otimes", "[ERROR], Otimes only accept rank-2 UniTensors as arguments.")
    if a.braket is not None or b.braket is not None:
        raise TypeError("linalg.Otimes", "Otimes can only accept untagged tensor.")
    if a.is_diag and b.is_diag:
        return _rank2(torch.outer(a._dense, b._dense).reshape(-1), is_diag=True)
    ta = torch.diag(a._dense) if a.is_diag else a._dense
    tb = torch.diag(b._dense) if b.is_diag else b._dense
    out = torch.tensordot(ta, tb, dims=0).permute(0, 2, 1, 3).reshape(ta.shape[0] * tb.shape[0], -1)
    return _rank2(out)


def ExpH(a):
    """exp of a Hermitian rank-2 tensor via eigh (linalg.py:260-319; torch.symeig no longer exists)."""
    if not isinstance(a, UniTensor):
        raise Exception("ExpH(UniTensor)", "[ERROR] ExpH can only accept UniTensor")
    if a.is_symm:
        raise Exception("ExpH(a)", "don't support symm tensor. GetBlock first.")
    if a.braket is not None:
        raise Exception("ExpH(a)", "can only accept [untagged] type UniTensor")
    if a.is_diag:
        return _rank2(torch.exp(a._dense), bonds=a.bonds, labels=a.labels, is_diag=True)
    if len(a.labels) != 2 or a.rowrank != 1:
        raise Exception("ExpH(a)", "a should be rank-2 tensor with 1 inbond 1 outbond")
    s, u = torch.linalg.eigh(a._dense, UPLO="U")
    s = torch.exp(s)
    return _rank2(torch.matmul(u * s, u.transpose(0, 1)), bonds=a.bonds, labels=a.labels)


def _check_rank2(a, who, allow_diag=False):
    if a.is_symm:
        raise Exception("%s(a)" % who, "[Abort] %s curretly don't support symm tensor." % who)
    if a.is_diag and not allow_diag:
        raise Exception("%s(a)" % who, "[Abort] %s currently don't support diagonal tensor." % who)
    if a.braket is not None:
        raise Exception("%s(a)" % who, "can only accept UniTensor[regular] (untagged)")
    if len(a.labels) != 2 or a.rowrank != 1:
        raise Exception("%s(a)" % who, "should be a UniTensor with 1 in-bond, 1 out-bond")


def Qr(a):
    if not isinstance(a, UniTensor):
        raise Exception("Qr(UniTensor)", "[ERROR] Qr can only accept UniTensor")
    _check_rank2(a, "Qr")
    q, r = torch.linalg.qr(a._dense)
    tmp = _new_label_ref(a)
    tq = _rank2(q, labels=[a.labels[0], tmp - 1])
    tr = _rank2(r, labels=[tq.labels[1], a.labels[1]])
    return tq, tr


def Qdr(a):
    if not isinstance(a, UniTensor):
        raise Exception("Qdr(UniTensor)", "[ERROR] Qdr can only accept UniTensor")
    _check_rank2(a, "Qdr")
    q, r = torch.linalg.qr(a._dense)
    d = r.diag()
    r = (r.t() / d).t()
    tmp = _new_label_ref(a)
    tq = _rank2(q, labels=[a.labels[0], tmp - 1])
    td = _rank2(d, labels=[tmp - 1, tmp - 2], is_diag=True)
    tr = _rank2(r, labels=[td.labels[1], a.labels[1]])
    return tq, td, tr


_SVD_SMALL = os.environ.get("TOR10_SVD_SMALL", "1") == "1"
_SVD_LIMITS = []
_SVD_WARM = os.environ.get("TOR10_SVD_WARM", "1") == "1"
_SVD_STATE = {}
_SVD_DEBUG = []          # [int32 device word, list]: set by tools/svd_sweeps.py


def _svd_small(x):
    """u, s, vh of a small 2-D CUDA matrix by the one-CTA Jacobi kernel (t10_svd_small: 64 x 64 in ~0.1 ms where
    cuSOLVER needs 1.8 ms), or None when the matrix does not fit one SM (the caller then uses torch / cuSOLVER)."""
    import ctypes as C
    on_gpu = x.is_cuda or getattr(_lib, "_EMU", False)       # (_EMU: the C-ABI emulator of the CPU test suite)
    if x.dim() != 2 or not on_gpu or x.dtype not in (torch.float64, torch.float32) or x.requires_grad:
        return None
    if not _SVD_LIMITS:
        mn, me = C.c_int(0), C.c_int(0)
        _lib.check(_lib.load().t10_svd_small_limits(C.byref(mn), C.byref(me)))
        _SVD_LIMITS.extend([int(mn.value), int(me.value)])
    m, n = int(x.shape[0]), int(x.shape[1])
    wide = m < n
    mm, nn = (n, m) if wide else (m, n)
    ne = (nn + 1) // 2 * 2
    if nn < 1 or nn > _SVD_LIMITS[0] or ((mm | 1) + (ne | 1)) * ne * x.element_size() > _SVD_LIMITS[1] * 8:
        return None
    rs, cs = int(x.stride(0)), int(x.stride(1))
    if wide:
        rs, cs = cs, rs
    u = torch.empty((mm, nn), device=x.device, dtype=x.dtype)
    s = torch.empty((nn,), device=x.device, dtype=x.dtype)
    vh = torch.empty((nn, nn), device=x.device, dtype=x.dtype)
    # warm-start state per (shape, dtype, device): consecutive SVDs of slowly changing matrices (iTEBD) start from
    # the previous right singular vectors (t10_svd_small)
    W = None
    if _SVD_WARM and ((mm | 1) * (ne + nn) + (ne | 1) * ne) * x.element_size() + ne * ne <= _SVD_LIMITS[1] * 8:
        key = (mm, nn, rs, cs, x.dtype, x.device.index)
        W = _SVD_STATE.get(key)
        if W is None:
            if len(_SVD_STATE) >= 64:
                _SVD_STATE.pop(next(iter(_SVD_STATE)))
            W = _SVD_STATE[key] = torch.zeros(2 * (nn * nn + 1) + 1, device=x.device, dtype=x.dtype)   # T10_SVD_WARM_ELEMS
    prob = (C.c_int64 * 9)(mm, nn, 0, rs, cs, 0, 0, 0, 0 if W is not None else -1)
    with _lib.on_device(x.device):
        _lib.check(_lib.load().t10_svd_small(_lib.dtype_code(x.dtype), 1, prob, x.data_ptr(), u.data_ptr(),
                                             s.data_ptr(), vh.data_ptr(), W.data_ptr() if W is not None else None,
                                             _SVD_DEBUG[0].data_ptr() if _SVD_DEBUG else None, _lib.stream_ptr()))
    if _SVD_DEBUG:          # tools/svd_sweeps.py: sweeps per call (synchronises)
        _SVD_DEBUG[1].append(int(_SVD_DEBUG[0].item()) >> 1)
    if wide:          # A^T = u s vh  =>  A = vh^T s u^T
        return vh.t(), s, u.t()
    return u, s, vh


def _svd(a, keepdim, who):
    _check_rank2(a, who)
    # torch.svd(some=True) of the reference == linalg.svd(full_matrices=False) with V = Vh^T.  Default cuSOLVER
    # driver = torch's (gesvdj with gesvd fallback).  TOR10_SVD_DRIVER=gesvda is 1.7x (64x64) to 8x (1024x1024)
    # faster on B200 (tools/svd_drivers.py) but fails to converge on rank-deficient sector blocks and shifts
    # the iTEBD energy trajectory by 3e-9, so it is opt-in.
    x = a._dense
    small = _svd_small(x) if _SVD_SMALL else None
    if small is not None:
        u, s, vh = small
    else:
        drv = os.environ.get("TOR10_SVD_DRIVER") if x.device.type == "cuda" else None
        u, s, vh = torch.linalg.svd(x, full_matrices=False, driver=drv or None)
    tmp = _new_label_ref(a)
    if keepdim is not None:
        if keepdim < 0 or keepdim > len(s):
            raise ValueError("Svd_truncate", "[ERROR] the keepdim=%d is invalid, must larger than 0 and smaller than "
                                             "the total number of eigenvalues." % keepdim)
        u, s, vh = u[:, :keepdim], s[:keepdim], vh[:keepdim, :]
    tu = _rank2(u, labels=[a.labels[0], tmp - 1])
    tv = _rank2(vh, labels=[tmp - 2, a.labels[1]])
    ts = UniTensor(bonds=[tu.bonds[1], tv.bonds[0]], labels=[tu.labels[1], tv.labels[0]], rowrank=1, check=False,
                   is_diag=True)
    ts._mac(torch_tensor=s)
    return tu, ts, tv


_SVD_POOL = {}


def _sector_svds(one, views, mine, device):
    """The independent sector SVDs of a symmetric tensor.  A cuSOLVER Jacobi SVD is a few hundred tiny kernels with a
    host poll per sweep: back to back on one stream the GPU idles most of the time (config 4: 34 sectors, 0.56 s).
    Sectors too large for the one-CTA kernel are therefore factorised CONCURRENTLY, one host thread + CUDA stream per
    worker (torch releases the GIL inside the call; cuSOLVER handles are per thread and stream), largest first; the
    streams are joined back into the current stream before anything is returned (no hidden streams)."""
    nwork = os.environ.get("TOR10_SVD_STREAMS")
    if nwork is None:
        # one host thread per concurrent SVD: share the box's cores between the ranks running on it (2 ranks x 16
        # threads on 16 cores measured 0.22 s against 0.096 s for one rank)
        local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
        nwork = max(2, min(16, (os.cpu_count() or 16) // max(local, 1)))
    nwork = int(nwork)
    fac = {}
    rest = []
    for i in mine:
        if _SVD_SMALL and _svd_small_fits(views[i]):
            fac[i] = one(views[i])
        else:
            rest.append(i)
    if len(rest) < 2 or nwork < 2 or device.type != "cuda":
        for i in rest:
            fac[i] = one(views[i])
        return fac
    import concurrent.futures as cf
    key = (device.index, nwork)
    pool = _SVD_POOL.get(key)
    if pool is None:
        pool = _SVD_POOL[key] = (cf.ThreadPoolExecutor(max_workers=nwork),
                                 [torch.cuda.Stream(device=device) for _ in range(nwork)])
    ex, streams = pool
    cur = torch.cuda.current_stream(device)
    ready = torch.cuda.Event()
    ready.record(cur)
    rest.sort(key=lambda i: -int(views[i].shape[0]) * int(views[i].shape[1]) * min(views[i].shape))
    chunks = [rest[w::nwork] for w in range(nwork)]

    def work(w):
        out = {}
        with torch.cuda.device(device), torch.cuda.stream(streams[w]):
            streams[w].wait_event(ready)
            for i in chunks[w]:
                out[i] = one(views[i])
            done = torch.cuda.Event()
            done.record(streams[w])
        return out, done
    for out, done in ex.map(work, range(nwork)):
        fac.update(out)
        cur.wait_event(done)
        for tup in out.values():
            for t in tup:
                t.record_stream(cur)
    return fac


def _svd_small_fits(x):
    if not _SVD_LIMITS:
        import ctypes as C
        mn, me = C.c_int(0), C.c_int(0)
        _lib.check(_lib.load().t10_svd_small_limits(C.byref(mn), C.byref(me)))
        _SVD_LIMITS.extend([int(mn.value), int(me.value)])
    m, n = int(x.shape[0]), int(x.shape[1])
    mm, nn = (n, m) if m < n else (m, n)
    ne = (nn + 1) // 2 * 2
    return 1 <= nn <= _SVD_LIMITS[0] and ((mm | 1) + (ne | 1)) * ne * x.element_size() <= _SVD_LIMITS[1] * 8


def _svd_sym(a, keepdim):
    """Sector-wise SVD of a symmetric tensor split as (row bonds | col bonds), with ONE global truncation to the
    `keepdim` largest singular values over all sectors (EXTENSION: the reference refuses symmetric tensors,
    linalg.py:630-631 / UniTensor.py:1672-1680).  Returns symmetric tensors
        U [row bonds ; x]     x = new BRA bond carrying the kept sector charges
        S [x' ; x]            block-diagonal, one diagonal block per kept sector
        V [x' ; col bonds]    x' = the same bond as KET
    so that Contract(Contract(U, S), V) reproduces (the truncation of) `a`.  The per-sector SVDs run on the
    device through torch/cuSOLVER; only the kept counts come back to the host (they size the new bond)."""
    from .Bond import BD_BRA, BD_KET
    from . import parallel
    shard = parallel.current_shard()
    sharded = shard is not None and parallel._STATE["transport"] in ("auto", "resident", "pull") \
        and parallel.transport_for() in ("resident", "pull")
    if sharded and a._res is not None and a._layout_current():
        src = a                                  # a sector-resident Contract result: every rank factorises what it owns
    else:
        src = a.Contiguous()
    L = src._layout
    drv = os.environ.get("TOR10_SVD_DRIVER") if src.device.type == "cuda" else None
    owner = None
    if sharded:
        # independent sector SVDs are dealt over the ranks: a resident operand keeps its owners, a
        # replicated one is split by LPT on the SVD cost m n min(m, n)
        rank, world = shard
        if src._res is not None:
            owner = np.asarray(src._res.owner).copy()
            free = owner < 0
            owner[free] = np.arange(int(free.sum())) % world        # zero sectors: anybody
        else:
            cost = L.rows.astype(np.float64) * L.cols * np.minimum(L.rows, L.cols)
            owner = parallel.partition_lpt(cost, world)
        views = L.views(src._arena_tensor(resident_ok=True))
        mine = [i for i in range(L.nsec) if owner[i] == rank]
    else:
        views = L.views(src._arena_tensor())
        mine = list(range(L.nsec))

    def one(v):
        small = _svd_small(v) if _SVD_SMALL else None
        return small if small is not None else torch.linalg.svd(v, full_matrices=False, driver=drv or None)
    fac = _sector_svds(one, views, mine, src.device)
    counts = [int(min(L.rows[i], L.cols[i])) for i in range(L.nsec)]
    total = sum(counts)
    if keepdim is None:
        keep = counts
    else:
        if keepdim <= 0 or keepdim > total:
            raise ValueError("Svd_truncate", "[ERROR] the keepdim=%d is invalid, must larger than 0 and smaller than "
                                             "the total number of eigenvalues." % keepdim)
        if sharded:
            # every rank contributes the singular values of its sectors; one small all-reduce completes the list
            import torch.distributed as dist
            starts = np.concatenate([[0], np.cumsum(counts)])
            all_s = torch.zeros(total, device=src.device, dtype=src.dtype)
            for i in mine:
                all_s[int(starts[i]):int(starts[i + 1])] = fac[i][1]
            dist.all_reduce(all_s, group=parallel._STATE["group"])
        else:
            all_s = torch.cat([fac[i][1] for i in range(L.nsec)])
        sec_id = torch.cat([torch.full((c,), i, dtype=torch.int64, device=all_s.device) for i, c in enumerate(counts)])
        top = torch.topk(all_s, keepdim).indices
        keep = torch.bincount(sec_id[top], minlength=len(counts)).cpu().tolist()   # prefix of every sector
    kept = [(i, k) for i, k in enumerate(keep) if k > 0]
    K = sum(k for _, k in kept)
    newq = np.concatenate([np.repeat(L.block_qnums[i][None, :], k, axis=0) for i, k in kept], axis=0)
    syms = src.bonds[0].sym_types
    rr = src.rowrank
    tmp = _new_label(a)
    dev, dt = src.device, src.dtype
    xb = Bond(K, BD_BRA, qnums=newq.tolist(), sym_types=syms)
    xk = Bond(K, BD_KET, qnums=newq.tolist(), sym_types=syms)
    U = UniTensor(bonds=list(src.bonds[:rr]) + [xb], rowrank=rr, labels=list(src.labels[:rr]) + [tmp - 1],
                  device=dev, dtype=dt)
    S = UniTensor(bonds=[xk, xb], rowrank=1, labels=[tmp - 1, tmp - 2], device=dev, dtype=dt)
    V = UniTensor(bonds=[xk] + list(src.bonds[rr:]), rowrank=1, labels=[tmp - 2] + list(src.labels[rr:]),
                  device=dev, dtype=dt)
    uv, sv, vv = (t._layout.views(t._arena) for t in (U, S, V))
    own = [np.full(t._layout.nsec, -1, dtype=np.int64) for t in (U, S, V)] if sharded else None
    for i, k in kept:
        q = L.block_qnums[i]
        if sharded:
            for j, t in enumerate((U, S, V)):
                own[j][t._layout.sector_of(q)] = owner[i]
        if i not in fac:
            continue
        u, s_, vh = fac[i]
        uv[U._layout.sector_of(q)].copy_(u[:, :k])
        sv[S._layout.sector_of(q)].copy_(torch.diag(s_[:k]))
        vv[V._layout.sector_of(q)].copy_(vh[:k, :])
    if sharded:
        # the factors stay sector-resident: every rank holds the sectors it factorised (not in peer-readable memory:
        # a consumer that needs foreign sectors completes them by a collective, parallel.gather_collective)
        for j, t in enumerate((U, S, V)):
            t._res = parallel.Residency(shard[1], shard[0], own[j], 0, None)
    return U, S, V


def Svd(a):
    if not isinstance(a, UniTensor):
        raise Exception("Svd(UniTensor)", "[ERROR] Svd can only accept UniTensor")
    if a.is_symm:
        return _svd_sym(a, None)
    return _svd(a, None, "svd")


def Svd_truncate(a, keepdim=None):
    if not isinstance(a, UniTensor):
        raise Exception("Svd_truncate(UniTensor,int)", "[ERROR] Svd_truncate can only accept UniTensor")
    if a.is_symm:
        return _svd_sym(a, keepdim)
    return _svd(a, keepdim, "svd")


def _matmul_sym(a, b):
    """Matmul of two symmetric rank-2 tensors (extension; the reference aborts, linalg.py:699-700):
    the sector-wise product on the grouped-GEMM kernel, i.e. Contract over a's column bond and b's row bond.
    Operands are not modified; the result carries the default labels [0, 1] like the dense Matmul."""
    import copy
    from .UniTensor import Contract
    if not (a.is_symm and b.is_symm):
        raise Exception("Matmul(a,b)", "[ERROR] cannot multiply a symmetric with a non-symmetric UniTensor")
    if not (len(a.labels) == 2 and len(b.labels) == 2 and a.rowrank == 1 and b.rowrank == 1):
        raise Exception("Matmul(a,b)", "Matmul can only accept two UniTensor with each has 1 inbond and 1 outbond")
    ta, tb = copy.copy(a), copy.copy(b)
    ta.labels = np.array([-1, -2], dtype=np.int64)
    tb.labels = np.array([-2, -3], dtype=np.int64)
    out = Contract(ta, tb)
    out.labels = np.array([0, 1], dtype=np.int64)
    return out


def Matmul(a, b):
    """linalg.py:680-737.  diag x diag: the reference returns the DOT PRODUCT of the two diagonals flagged
    is_diag (torch.matmul of two 1-D tensors, :717-722) -- a bug; the element-wise product (the diagonal
    of the matrix product) is returned here instead.  Symmetric operands: see _matmul_sym."""
    if not (isinstance(a, UniTensor) and isinstance(b, UniTensor)):
        raise TypeError("_Matmul(a,b)", "[ERROR] _Matmul can only accept UniTensors for both a & b")
    if a.is_symm or b.is_symm:
        return _matmul_sym(a, b)
    if a.braket is not None or b.braket is not None:
        raise Exception("Matmul(a,b)", "Matmul can only accept two regular Tensors")
    if not (len(a.labels) == 2 and len(b.labels) == 2 and a.rowrank == 1 and b.rowrank == 1):
        raise Exception("Matmul(a,b)", "Matmul can only accept two UniTensor with each has 1 inbond and 1 outbond")
    if int(a.bonds[1].dim) != int(b.bonds[0].dim):
        # the reference leaves this to torch.matmul (linalg.py:722-733), which raises RuntimeError
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)"
                           % (a.bonds[0].dim, a.bonds[1].dim, b.bonds[0].dim, b.bonds[1].dim))
    bonds = [a.bonds[0], b.bonds[1]]
    if a.is_diag and b.is_diag:
        return _rank2(a._dense * b._dense, bonds=bonds, is_diag=True)
    if a.is_diag:
        return _rank2(_scale(b._dense, a._dense, side=0), bonds=bonds)
    if b.is_diag:
        return _rank2(_scale(a._dense, b._dense, side=1), bonds=bonds)
    return _rank2(_dense_gemm(UniTensor._contiguous_dense(a._dense), UniTensor._contiguous_dense(b._dense)),
                  bonds=bonds)


def _scale(mat, diag, side):
    """diag(d) @ mat (side 0) or mat @ diag(d) (side 1): family-2 scaling kernel, no densified diagonal."""
    from .UniTensor import _diag_scale
    return _diag_scale(mat, diag, 0 if side == 0 else 1)


def Chain_matmul(*args):
    """linalg.py:740-810: left-to-right chain on the GEMM kernel."""
    if not all(isinstance(x, UniTensor) for x in args):
        raise TypeError("_Chain_matmul(*args)", "[ERROR] _Chain_matmul can only accept UniTensors for all elements in "
                                                "args")
    if not all(len(x.labels) == 2 and x.rowrank == 1 for x in args):
        raise Exception("Chain_matmul(*args)", "Chain mult should have all UniTensor have 1 inbond 1 outbond")
    out = args[0]
    for nxt in args[1:]:
        out = Matmul(out, nxt)
    if out.is_symm:      # extension (the reference aborts, linalg.py:764-765): chain of sector-wise products
        return out
    res = UniTensor(bonds=[args[0].bonds[0], args[-1].bonds[1]], rowrank=1, check=False)
    res._mac(torch_tensor=torch.diag(out._dense) if out.is_diag else out._dense)
    return res


def Inverse(a):
    if not isinstance(a, UniTensor):
        raise Exception("Inverse(UniTensor)", "[ERROR] Inverse can only accept UniTensor")
    if a.is_symm:
        raise TypeError("Inverse", "[ERROR] cannot inverse a symmetry tensor")
    if a.braket is not None:
        raise TypeError("Inverse", "[ERROR] inverse can only accept untagged UniTensor[regular]")
    if a.rowrank != 1 or len(a.labels) != 2:
        raise Exception("Inverse", "[ERROR] inverse should have UniTensor with rowrank=1")
    if a.is_diag:
        return _rank2(a._dense ** -1, bonds=a.bonds, labels=a.labels, is_diag=True)
    if a._diag_hint is not None:
        # a dense matrix known to be diag(d) (UniTensor._dense): its inverse is diag(1 / d) -- what the LU of
        # torch.inverse computes for such a matrix, without the factorisation
        r = a._diag_hint ** -1
        out = _rank2(torch.diag(r), bonds=a.bonds, labels=a.labels)
        out.__dict__["_diag_hint"] = r
        return out
    return _rank2(torch.inverse(a._dense), bonds=a.bonds, labels=a.labels)


def Det(a):
    if not isinstance(a, UniTensor):
        raise Exception("Det(UniTensor)", "[ERROR] Det can only accept UniTensor")
    if a.is_symm:
        raise Exception("Det", "[ERROR] cannot operate deteminant on a symmetry tensor")
    if a.braket is not None:
        raise Exception("Det", "[ERROR] det can only operate on untagged UniTensor[regular]")
    if a.rowrank != 1 or len(a.labels) != 2:
        raise Exception("Det", "[ERROR] det should have a rank-2 UniTensor with rowrank=1")
    return _scalar(torch.prod(a._dense) if a.is_diag else torch.det(a._dense))


def Norm(a):
    if not isinstance(a, UniTensor):
        raise Exception("Norm(UniTensor)", "[ERROR] Norm can only accept UniTensor")
    if a.is_symm:
        raise Exception("Norm", "[ERROR] cannot operate Norm on a symmetry tensor")
    if a.braket is not None:
        raise Exception("Norm", "[ERROR] Norm can only operate on untagged UniTensor")
    return _scalar(torch.norm(a._dense))
