"""TEST INFRASTRUCTURE -- never imported by the package.

A numpy stand-in for libtor10_b200.so: every C-ABI entry point the Python host calls is restated here on host
memory (block maps through the oracle, relayouts and GEMMs as numpy expressions of what include/tor10_b200.h
documents).  Importing this module swaps it in for `tor10_b200._lib.load()` IN THE CURRENT PROCESS, so that the
host logic of the package -- plans, label algebra, metadata, error behaviour, the stream-K / tile planners (those
are host code of the real library and are called for real) -- can be exercised on a machine without a GPU:

  * tests/test_host_emulated.py runs the GPU parity suites against it in a SUBPROCESS (pytest -p emu_plugin);
  * tests/fuzz_vs_reference.py compares the package, through it, with the unmodified reference on random inputs.

It proves nothing about the CUDA kernels (the `-m gpu` tests do, through the real library), and the product has
no way to reach it: without CUDA the package raises (tests/test_host_cpu.py::test_no_cpu_compute_path).
"""
import ctypes as C, contextlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tor10_oracle as orc
import tor10_b200 as T
from tor10_b200 import _lib, _plan

real = _lib.load()
_KEEP = []
def arr(ptr, n, dt):
    if isinstance(ptr, C.Array):
        _KEEP.append(ptr); ptr = C.addressof(ptr)
    if hasattr(ptr, "value"):
        _KEEP.append(ptr)
        if len(_KEEP) > 4096: del _KEEP[:2048]
        ptr = ptr.value
    if n == 0: return np.zeros(0, dt)
    ct = {np.int64: C.c_int64, np.int32: C.c_int32, np.float64: C.c_double, np.float32: C.c_float, np.uint8: C.c_uint8}[dt]
    return np.ctypeslib.as_array((ct * n).from_address(ptr))

class Emu:
    def t10_last_error(self): return b"emu"
    def t10_version(self): return 1
    def t10_ggemm_plan_tiles(self, *a): return real.t10_ggemm_plan_tiles(*a)
    def t10_ggemm_tile_shape(self, *a): return real.t10_ggemm_tile_shape(*a)
    def t10_fuse_qnums(self, nb, nsym, dims, signs, kind, symn, qoff, dq, dout, st):
        dims = arr(dims, nb, np.int64); signs = arr(signs, nb, np.int32); kind = arr(kind, nsym, np.int32); symn = arr(symn, nsym, np.int64); qoff = arr(qoff, nb, np.int64)
        syms = [("U1", 0) if kind[s] == 0 else ("Zn", int(symn[s])) for s in range(nsym)]
        tot = int(dims.sum())
        q = arr(dq, tot * nsym, np.int64).reshape(tot, nsym)
        v = q[qoff[0]:qoff[0] + dims[0]] * int(signs[0])
        for b in range(1, nb):
            v = orc.combine_qnums(v, q[qoff[b]:qoff[b] + dims[b]] * int(signs[b]), syms)
        arr(dout, v.size, np.int64)[:] = v.reshape(-1)
        return 0
    def t10_blockmap_build(self, rank, rr, nsym, dims, braket, kind, symn, qoff, dq, cap, bq, km, bm, ki, bi, si, rb, co, counts, st):
        dims = arr(dims, rank, np.int64); braket = arr(braket, rank, np.int32); kind = arr(kind, nsym, np.int32); symn = arr(symn, nsym, np.int64); qoff = arr(qoff, rank, np.int64)
        syms = [("U1", 0) if kind[s] == 0 else ("Zn", int(symn[s])) for s in range(nsym)]
        tot = int(dims.sum()); q = arr(dq, tot * nsym, np.int64).reshape(tot, nsym)
        bonds = [orc.OBond(int(dims[b]), int(braket[b]), q[qoff[b]:qoff[b] + dims[b]], syms, _presorted=True) for b in range(rank)]
        try:
            m = orc.build_blockmap(bonds, braket.astype(np.int64), rr)
        except TypeError:
            return 4
        R, Cc = len(m.ket_map), len(m.bra_map); ns = len(m.block_qnums)
        arr(bq, cap * nsym, np.int64).reshape(cap, nsym)[:ns] = m.block_qnums
        arr(km, R * 2, np.int64)[:] = m.ket_map.reshape(-1); arr(bm, Cc * 2, np.int64)[:] = m.bra_map.reshape(-1)
        kidx = np.concatenate(m.idx_in); bidx = np.concatenate(m.idx_out)
        arr(ki, R, np.int64)[:len(kidx)] = kidx; arr(bi, Cc, np.int64)[:len(bidx)] = bidx
        info = arr(si, cap * 8, np.int64).reshape(cap, 8); info[:] = 0
        off = 0; ks = 0; bs = 0
        rbv = arr(rb, R, np.int64); rbv[:] = -1; cov = arr(co, Cc, np.int32); cov[:] = -1
        for s, (r, c) in enumerate(m.shapes):
            ld = c + (c & 1)
            info[s, :6] = (r, c, ks, bs, off, ld)
            rbv[m.idx_in[s]] = off + np.arange(r) * ld
            cov[m.idx_out[s]] = np.arange(c)
            ks += r; bs += c; off += -(-(r * ld) // 16) * 16
        arr(counts, 4, np.int64)[:] = (ns, off, ks, bs)
        return 0
    def t10_unravel(self, n, ns, strides, idx, out, st):
        strides = arr(strides, ns, np.int64); x = arr(idx, n, np.int64).copy(); o = arr(out, n * ns, np.int64).reshape(n, ns)
        for p in range(ns):
            o[:, p] = x // strides[p]; x = x % strides[p]
        return 0
    def t10_relayout_tables(self, n, nb, dims, srow, scol, idx, tr, tc, st):
        dims = arr(dims, nb, np.int64); srow = arr(srow, nb, np.int64); scol = arr(scol, nb, np.int64)
        x = arr(idx, n, np.int64).copy(); r = np.zeros(n, np.int64); c = np.zeros(n, np.int64)
        for b in range(nb - 1, -1, -1):
            d = x % dims[b]; x //= dims[b]; r += d * srow[b]; c += d * scol[b]
        arr(tr, n, np.int32)[:] = r; arr(tc, n, np.int32)[:] = c
        return 0
    def t10_relayout_index(self, dt, src, dst, n, index, st):
        fd = np.float64 if dt == 0 else np.float32
        ix = arr(index, n, np.int32); s_ = arr(src, self.ctx['src_n'], fd); d_ = arr(dst, n, fd)
        m = ix >= 0; d_[m] = s_[ix[m]]; d_[ix == -1] = 0
        return 0
    def t10_reduce_slices(self, dt, ws, out, n, ns, st):
        fd = np.float64 if dt == 0 else np.float32
        w = arr(ws, n * ns, fd).reshape(ns, n); o = arr(out, n, fd)
        acc = w[0].copy()
        for s_ in range(1, ns): acc += w[s_]
        o[:] = acc
        return 0
    def t10_device_props(self, dev, nsm, a, b, c):
        C.cast(nsm, C.POINTER(C.c_int))[0] = 148
        return 0
    def t10_relayout_runs(self, dt, src, dst, n, runs, nruns, chunk_first, st):
        fd = np.float64 if dt == 0 else np.float32
        rr = arr(runs, (nruns + 1) * 2, np.int32).reshape(-1, 2).astype(np.int64)
        s_ = arr(src, self.ctx['src_n'], fd); d_ = arr(dst, n, fd)
        cf = arr(chunk_first, (n + 1023) // 1024 + 1, np.int32)
        assert cf[-1] == nruns - 1 and rr[nruns, 0] == n
        for r in range(nruns):
            a, b, so = rr[r, 0], rr[r + 1, 0], rr[r, 1]
            assert cf[a // 1024] <= r
            if so >= 0: d_[a:b] = s_[so:so + (b - a)]
            elif so == -1: d_[a:b] = 0
        return 0
    def t10_relayout_blocks(self, dt, src, dst, ntiles, tiles, RR, RC, CR, CC, rowbase, coloff, km, bm, emit, st):
        ctx = self.ctx
        fd = np.float64 if dt == 0 else np.float32
        tiles = arr(tiles, ntiles * 32, np.uint8).view(_lib.RELAYOUT_TILE_DTYPE)
        RR = arr(RR, ctx["nket"], np.int32); RC = arr(RC, ctx["nket"], np.int32); CR = arr(CR, ctx["nbra"], np.int32); CC = arr(CC, ctx["nbra"], np.int32)
        rowbase = arr(rowbase, ctx["R_M"], np.int64); coloff = arr(coloff, ctx["C_M"], np.int32)
        if emit:
            d_ = arr(emit, ctx['dst_n'] if 'dst_n' in ctx else 0, np.int32); s_ = None
        else:
            s_ = arr(src, ctx["src_n"], fd); d_ = arr(dst, ctx["dst_n"], fd)
        KM = arr(km, ctx['R_M']*2, np.int64) if km else None; BM = arr(bm, ctx['C_M']*2, np.int64) if bm else None
        for tl in tiles:
            for r in range(int(tl["nr"])):
                for c in range(int(tl["nc"])):
                    sr = RR[tl["rbase"] + r] + CR[tl["cbase"] + c]; sc = RC[tl["rbase"] + r] + CC[tl["cbase"] + c]
                    base = rowbase[sr]; co = coloff[sc]
                    ok = base >= 0 and co >= 0
                    if ok and KM is not None: ok = KM[2*sr] == BM[2*sc]
                    if emit: d_[int(tl["dst"]) + r * int(tl["dld"]) + c] = (base + co) if ok else -1
                    else: d_[int(tl["dst"]) + r * int(tl["dld"]) + c] = s_[base + co] if ok else 0
        return 0
    def t10_permute_dense(self, dt, src, dst, rank, shape, strides, st):
        fd = np.float64 if dt == 0 else np.float32
        shape = arr(shape, rank, np.int64); strides = arr(strides, rank, np.int64)
        n = int(np.prod(shape)); span = int(((shape - 1) * strides).sum()) + 1
        s_ = arr(src, span, fd)
        v = np.lib.stride_tricks.as_strided(s_, shape=tuple(shape), strides=tuple(strides * s_.itemsize))
        arr(dst, n, fd)[:] = np.ascontiguousarray(v).reshape(-1)
        return 0
    def t10_diag_scale(self, dt, a, d, out, outer, dd, inner, st):
        fd = np.float64 if dt == 0 else np.float32
        A = arr(a, outer * dd * inner, fd).reshape(outer, dd, inner); D = arr(d, dd, fd)
        arr(out, outer * dd * inner, fd).reshape(outer, dd, inner)[:] = A * D[None, :, None]
        return 0
    def t10_ggemm_tf32_status(self, out):
        out._obj.value = 0
        return 0
    def t10_ggemm_streamk(self, A, B, Cc, nprob, prob, npieces, tiles, pieces, nslots, slots, ws, flags, epoch, st):
        return self.t10_ggemm(0, 46, A, B, Cc, nprob, prob, npieces, tiles, nslots, slots, None, st)
    def t10_relayout_segments(self, dt, nsrc, hsrc, dst, nseg, seg, st):
        fd = np.float64 if dt == 0 else np.float32
        sg = arr(seg, nseg * 4, np.int32).reshape(nseg, 4)
        bases = [int(x or 0) for x in hsrc]
        isz = np.dtype(fd).itemsize
        for d, s_, ln, w in sg.tolist():
            o = arr(dst + d * isz, ln, fd)
            o[:] = 0 if s_ < 0 else arr(bases[w] + s_ * isz, ln, fd)
        return 0
    def t10_svd_small_limits(self, mn, me):
        mn._obj.value = 64; me._obj.value = 25600
        return 0
    def t10_svd_small(self, dt, nprob, prob, A, U, S, Vh, W, info, st):
        fd = np.float64 if dt == 0 else np.float32
        isz = np.dtype(fd).itemsize
        pr = arr(prob, 9 * nprob, np.int64).reshape(nprob, 9)
        for m, n, ao, rs, cs, uo, so, vo, wo in pr.tolist():
            span = (m - 1) * rs + (n - 1) * cs + 1
            a = np.lib.stride_tricks.as_strided(arr(A + ao * isz, span, fd), (m, n), (rs * isz, cs * isz))
            u, s_, vh = np.linalg.svd(a, full_matrices=False)
            arr(U + uo * isz, m * n, fd)[:] = u.reshape(-1); arr(S + so * isz, n, fd)[:] = s_; arr(Vh + vo * isz, n * n, fd)[:] = vh.reshape(-1)
        return 0
    def t10_ggemm_tma_slots(self, out):
        out._obj.value = 296
        return 0
    def t10_ggemm_tma_encode(self, nprob, prob, A, B, hmaps):
        arr(hmaps, 2, np.int64)[:] = [A or 0, B or 0]       # the emulated "maps" are just the operand bases
        return 0
    def t10_ggemm_tma(self, maps, Cc, nprob, prob, npieces, tiles, pieces, nslots, slots, ws, flags, epoch, pdl, nsig, sigp, sigv, done, npush, push, st):
        A, B = (int(x) for x in arr(maps, 2, np.int64))
        return self.t10_ggemm(0, 46, A, B, Cc, nprob, prob, npieces, tiles, nslots, slots, None, st)
    def t10_ggemm_slots(self, dt, cfg, out):
        out._obj.value = 296
        return 0
    def t10_ggemm(self, dt, cfg, A, B, Cc, nprob, prob, ntiles, tiles, nslots, slots, sched, st):
        fd = np.float64 if dt == 0 else np.float32
        p = arr(prob, nprob * 48, np.uint8).view(_lib.GEMM_PROBLEM_DTYPE)
        for q in p:
            m, n, k, lda, ldb, ldc = int(q["m"]), int(q["n"]), int(q["k"]), int(q["lda"]), int(q["ldb"]), int(q["ldc"])
            a = arr(A + int(q["a_off"]) * np.dtype(fd).itemsize, max(0, (m - 1) * lda + k), fd)
            b = arr(B + int(q["b_off"]) * np.dtype(fd).itemsize, max(0, (k - 1) * ldb + n), fd)
            c = arr(Cc + int(q["c_off"]) * np.dtype(fd).itemsize, max(0, (m - 1) * ldc + n), fd)
            am = np.lib.stride_tricks.as_strided(a, (m, k), (lda * a.itemsize, a.itemsize))
            bm = np.lib.stride_tricks.as_strided(b, (k, n), (ldb * b.itemsize, b.itemsize))
            cm = np.lib.stride_tricks.as_strided(c, (m, n), (ldc * c.itemsize, c.itemsize))
            cm[:] = am @ bm
        return 0

emu = Emu()
_lib.load = lambda: emu
_lib._EMU = True
_lib.require_cuda = lambda d: torch.device(d)
_lib.stream_ptr = lambda: 0
class _NoDev:
    def __init__(self, d): pass
    def __enter__(self): pass
    def __exit__(self, *a): return False
_lib.on_device = _NoDev
_lib.require_cuda_tensor = lambda x, w: None
_plan._dev_index = lambda d: 0
_plan._sm_count = lambda: 148
torch.cuda.device = lambda d=None: contextlib.nullcontext()
_orig_run = _plan.RelayoutPlan.run
def run(self, src, dst):
    emu.ctx = dict(nsec=self.L.nsec,  nket=self.L.n_ket, nbra=self.L.n_bra, R_M=self.M.R, C_M=self.M.C, src_n=src.numel(), dst_n=dst.numel())
    return _orig_run(self, src, dst)
_plan.RelayoutPlan.run = run
