"""Plans: everything about a symmetric tensor layout / a relayout / a contraction that depends only
on bond signatures, labels and row ranks -- computed once (device kernels of family 1), cached,
and then replayed with zero host bookkeeping per call.

  BlockLayout    block map of (bonds, braket, rowrank)         <- UniTensor.__init__ :393-441
  RelayoutPlan   address tables + tile list M -> L              <- UniTensor.Contiguous :2097-2129
  ContractPlan   relayouts + sector matching + GEMM tile table  <- Contract :3686-3743

The reference recomputes all of this with numpy on every call (3 block-map builds, 2 deepcopies and
an O(nsec^2) sector search per Contract).
"""
import ctypes as C
import itertools
import os

import numpy as np
import torch

from . import _lib
from ._lib import ARENA_ALIGN, RELAYOUT_TC, RELAYOUT_TR, check, i32, i64, np_ptr

class LRU:
    """Bounded plan cache: a dict with least-recently-used eviction on the number of entries and, when the
    entries report `plan_bytes()`, on the bytes of device tables they keep alive.  An evicted plan's tensors
    are released as soon as no other plan references them (a workload whose charge tables change every step --
    a symmetric Svd_truncate producing a new bond per iTEBD / DMRG step -- would otherwise grow until OOM)."""

    def __init__(self, max_entries, max_bytes=None):
        import collections
        self.d = collections.OrderedDict()
        self.max_entries, self.max_bytes = int(max_entries), max_bytes
        self.bytes = 0
        self.evicted = 0

    @staticmethod
    def _nbytes(v):
        f = getattr(v, "plan_bytes", None)
        return int(f()) if f is not None else 0

    def get(self, key, default=None):
        v = self.d.get(key)
        if v is None:
            return default
        self.d.move_to_end(key)
        return v

    def __setitem__(self, key, v):
        old = self.d.pop(key, None)
        if old is not None:
            self.bytes -= self._nbytes(old)
        self.d[key] = v
        self.bytes += self._nbytes(v)
        while len(self.d) > 1 and (len(self.d) > self.max_entries
                                   or (self.max_bytes is not None and self.bytes > self.max_bytes)):
            _, ev = self.d.popitem(last=False)
            self.bytes -= self._nbytes(ev)
            self.evicted += 1

    def recount(self):
        """Plan memory can grow after insertion (the element index is built on first use)."""
        self.bytes = sum(self._nbytes(v) for v in self.d.values())

    def __len__(self):
        return len(self.d)

    def __contains__(self, key):
        return key in self.d

    def items(self):
        return self.d.items()

    def values(self):
        return self.d.values()

    def clear(self):
        self.d.clear()
        self.bytes = 0


_CACHE_ENTRIES = int(os.environ.get("TOR10_PLAN_CACHE_ENTRIES", "512"))
_CACHE_BYTES = int(os.environ.get("TOR10_PLAN_CACHE_MB", "4096")) << 20
_LAYOUTS = LRU(_CACHE_ENTRIES, _CACHE_BYTES)
_RELAYOUTS = LRU(_CACHE_ENTRIES, _CACHE_BYTES)
_CONTRACTS = LRU(_CACHE_ENTRIES, _CACHE_BYTES)
STATS = {"layout_build": 0, "relayout_build": 0, "contract_build": 0}


def clear_caches():
    _LAYOUTS.clear()
    _RELAYOUTS.clear()
    _CONTRACTS.clear()
    import sys
    mod = sys.modules.get(__package__ + ".UniTensor")
    if mod is not None:
        mod._GEMM_GROUPS.clear()
    par = sys.modules.get(__package__ + ".parallel")
    if par is not None:
        par._SYMM.clear()
        par._STAGE.clear()


def pinned_objects():
    """Strong references to everything the caches hold right now: plans, relayouts, layouts, dense GEMM groups, the
    warm state of the small SVD.  A captured CUDA graph (graph.StaticLoop) addresses their device tables by raw
    pointer; holding this list keeps them alive after an LRU eviction or clear_caches()."""
    import sys
    pins = list(_LAYOUTS.values()) + list(_RELAYOUTS.values()) + list(_CONTRACTS.values())
    mod = sys.modules.get(__package__ + ".UniTensor")
    if mod is not None:
        pins += list(mod._GEMM_GROUPS.values())
    la = sys.modules.get(__package__ + ".linalg")
    if la is not None:
        pins += list(la._SVD_STATE.values())
    return pins


def _tensor_bytes(*ts):
    return sum(int(t.numel()) * t.element_size() for t in ts if t is not None)


def _owned_tensor_bytes(obj):
    """Bytes of the torch tensors an object holds directly or in a dict / list / tuple attribute (its device tables)."""
    n = 0
    for v in vars(obj).values():
        if isinstance(v, torch.Tensor):
            n += int(v.numel()) * v.element_size()
        elif isinstance(v, dict):
            n += sum(int(x.numel()) * x.element_size() for x in v.values() if isinstance(x, torch.Tensor))
        elif isinstance(v, (list, tuple)):
            n += sum(int(x.numel()) * x.element_size() for x in v if isinstance(x, torch.Tensor))
    return n


def _dev_index(device):
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


def _sym_codes(bond):
    kinds = i32([s._code()[0] for s in bond.sym_types])
    ns = i64([s._code()[1] for s in bond.sym_types])
    return kinds, ns


def accu_offsets(dims, rowrank):
    """UniTensor.py:396-403"""
    accu = np.ones(len(dims), dtype=np.int64)
    for i in range(len(dims) - 2, -1, -1):
        accu[i] = accu[i + 1] * int(dims[i + 1])
    accu_in = (accu[:rowrank] // accu[rowrank - 1]).astype(np.int64)
    accu_out = accu[rowrank:].astype(np.int64)
    return accu_in, accu_out


class BlockLayout:
    """Device-resident block map of one (bonds, braket, rowrank) signature."""

    def __init__(self, bonds, braket, rowrank, device):
        lib = _lib.load()
        device = _lib.require_cuda(device)
        STATS["layout_build"] += 1
        self.device = device
        self.rank = len(bonds)
        self.rowrank = int(rowrank)
        self.nsym = int(bonds[0].nsym)
        self.dims = i64([b.dim for b in bonds])
        self.braket = i32(braket)
        self.braket_key = np.asarray(braket, dtype=np.int64).tobytes()   # cheap equality test (UniTensor._layout_current)
        self.accu_in, self.accu_out = accu_offsets(self.dims, self.rowrank)
        self.R = int(np.prod(self.dims[:rowrank], dtype=np.int64))
        self.C = int(np.prod(self.dims[rowrank:], dtype=np.int64))
        kinds, ns = _sym_codes(bonds[0])
        self.all_u1 = bool((kinds == 0).all())
        qoff = np.zeros(self.rank, dtype=np.int64)
        qoff[1:] = np.cumsum(self.dims[:-1])
        qcat = np.concatenate([b.qnums for b in bonds], axis=0)
        with torch.cuda.device(device):
            d_q = torch.from_numpy(np.ascontiguousarray(qcat)).to(device)
            cap = min(self.R, self.C)
            opt = dict(device=device, dtype=torch.int64)
            self.d_block_qnums = torch.empty((cap, self.nsym), **opt)
            self.d_ket_map = torch.empty((self.R, 2), **opt)
            self.d_bra_map = torch.empty((self.C, 2), **opt)
            self.d_ket_idx = torch.empty((self.R,), **opt)
            self.d_bra_idx = torch.empty((self.C,), **opt)
            self.d_sec_info = torch.empty((cap, 8), **opt)
            self.d_rowbase = torch.empty((self.R,), **opt)
            self.d_coloff = torch.empty((self.C,), device=device, dtype=torch.int32)
            counts = np.zeros(4, dtype=np.int64)
            check(lib.t10_blockmap_build(
                self.rank, self.rowrank, self.nsym, np_ptr(self.dims), np_ptr(self.braket), np_ptr(kinds),
                np_ptr(ns), np_ptr(qoff), d_q.data_ptr(), cap, self.d_block_qnums.data_ptr(),
                self.d_ket_map.data_ptr(), self.d_bra_map.data_ptr(), self.d_ket_idx.data_ptr(),
                self.d_bra_idx.data_ptr(), self.d_sec_info.data_ptr(), self.d_rowbase.data_ptr(),
                self.d_coloff.data_ptr(), np_ptr(counts), _lib.stream_ptr()))
        self.nsec = int(counts[0])
        self.arena_elems = int(counts[1])
        self.n_ket = int(counts[2])
        self.n_bra = int(counts[3])
        self.d_block_qnums = self.d_block_qnums[:self.nsec]
        self.d_sec_info = self.d_sec_info[:self.nsec]
        self.d_ket_idx = self.d_ket_idx[:self.n_ket]
        self.d_bra_idx = self.d_bra_idx[:self.n_bra]
        self.block_qnums = self.d_block_qnums.cpu().numpy()            # [nsec, nsym] ascending lex
        info = self.d_sec_info.cpu().numpy()
        self.rows = info[:, 0].copy()
        self.cols = info[:, 1].copy()
        self.ket_start = info[:, 2].copy()
        self.bra_start = info[:, 3].copy()
        self.arena_off = info[:, 4].copy()
        self.ld = info[:, 5].copy()          # leading dimension of the stored sector = cols rounded up to even
        self.nnz = int((self.rows * self.cols).sum())
        self._qindex = {tuple(q): i for i, q in enumerate(self.block_qnums.tolist())}
        self._host = {}

    # ---- host mirrors in the reference's attribute format (lazy) -------------------
    def host(self, name):
        if name not in self._host:
            if name == "ket_map":
                v = self.d_ket_map.cpu().numpy()
            elif name == "bra_map":
                v = self.d_bra_map.cpu().numpy()
            elif name in ("ket_inv", "bra_inv"):
                idx = self.d_ket_idx if name == "ket_inv" else self.d_bra_idx
                strides = self.accu_in if name == "ket_inv" else self.accu_out
                start = self.ket_start if name == "ket_inv" else self.bra_start
                cnt = self.rows if name == "ket_inv" else self.cols
                out = torch.empty((idx.numel(), len(strides)), device=self.device, dtype=torch.int64)
                with torch.cuda.device(self.device):
                    check(_lib.load().t10_unravel(idx.numel(), len(strides), np_ptr(i64(strides)), idx.data_ptr(),
                                                  out.data_ptr(), _lib.stream_ptr()))
                h = out.cpu().numpy()
                v = [h[int(start[b]):int(start[b] + cnt[b])] for b in range(self.nsec)]
            else:
                raise KeyError(name)
            self._host[name] = v
        return self._host[name]

    def sector_of(self, qnum):
        return self._qindex.get(tuple(int(x) for x in qnum))

    def plan_bytes(self):
        return _tensor_bytes(self.d_block_qnums, self.d_ket_map, self.d_bra_map, self.d_ket_idx, self.d_bra_idx,
                             self.d_sec_info, self.d_rowbase, self.d_coloff)

    def new_arena(self, dtype, zero=True):
        n = self.arena_elems + ARENA_ALIGN
        return (torch.zeros if zero else torch.empty)(n, device=self.device, dtype=dtype)

    def views(self, arena):
        # [rows, cols] window of the [rows, ld] stored block (a strided view when cols is odd)
        return [arena[int(o):int(o + r * l)].view(int(r), int(l))[:, :int(c)]
                for o, r, c, l in zip(self.arena_off, self.rows, self.cols, self.ld)]


def layout_signature(bonds, braket, rowrank):
    return (tuple(b.signature()[0:1] + b.signature()[2:] for b in bonds), tuple(int(x) for x in braket), int(rowrank))


def get_layout(bonds, braket, rowrank, device):
    key = (layout_signature(bonds, braket, rowrank), _dev_index(device))
    lay = _LAYOUTS.get(key)
    if lay is None:
        lay = BlockLayout(bonds, braket, rowrank, device)
        lay.key = key
        _LAYOUTS[key] = lay
    return lay


# ------------------------------------------------------------------------------------
class PackedLayout:
    """Where a relayout writes: sector offsets / leading dimensions of the destination buffer.
    `native(L)` is the tensor's own arena; `padded(L)` is the GEMM-operand scratch with even
    leading dimensions (16-byte rows for the cp.async loader)."""

    def __init__(self, off, ld, size):
        self.off, self.ld, self.size = i64(off), i64(ld), int(size)

    @staticmethod
    def native(L):
        return PackedLayout(L.arena_off, L.ld, L.arena_elems + ARENA_ALIGN)   # multiple of 16

    @staticmethod
    def padded(L, sectors=None):
        """Scratch layout; with `sectors` only those sectors get space (a rank of a sharded contraction relays
        out 1/world of the tensor: the buffer and the element index shrink accordingly)."""
        ld = (L.cols + 1) // 2 * 2
        sz = (L.rows * ld + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
        if sectors is not None:
            keep = np.zeros(L.nsec, dtype=bool)
            keep[list(sectors)] = True
            sz = np.where(keep, sz, 0)
        off = np.zeros(L.nsec, dtype=np.int64)
        off[1:] = np.cumsum(sz)[:-1]
        return PackedLayout(off, ld, int(sz.sum()) + ARENA_ALIGN)


class RelayoutPlan:
    """src layout M --(mapper)--> dst layout L, optionally restricted to some dst sectors."""

    def __init__(self, M, L, mapper, packed, sectors=None):
        lib = _lib.load()
        STATS["relayout_build"] += 1
        self.M, self.L, self.packed = M, L, packed
        self.check = not M.all_u1      # U1 charges are conserved exactly; Zn layouts can lose sectors (Q3)
        dev = L.device
        mapper = [int(x) for x in mapper]
        rank = L.rank
        # logical position p sits at memory position mapper[p]; its digit moves the source flat row
        # (if that memory position is in M's row space) or the source flat column.
        srow = np.zeros(rank, dtype=np.int64)
        scol = np.zeros(rank, dtype=np.int64)
        for p in range(rank):
            mp = mapper[p]
            if mp < M.rowrank:
                srow[p] = M.accu_in[mp]
            else:
                scol[p] = M.accu_out[mp - M.rowrank]
        rr = L.rowrank
        o32 = dict(device=dev, dtype=torch.int32)
        self.RR = torch.empty(L.n_ket, **o32)
        self.RC = torch.empty(L.n_ket, **o32)
        self.CR = torch.empty(L.n_bra, **o32)
        self.CC = torch.empty(L.n_bra, **o32)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            check(lib.t10_relayout_tables(L.n_ket, rr, np_ptr(i64(L.dims[:rr])), np_ptr(i64(srow[:rr])),
                                          np_ptr(i64(scol[:rr])), L.d_ket_idx.data_ptr(), self.RR.data_ptr(),
                                          self.RC.data_ptr(), st))
            check(lib.t10_relayout_tables(L.n_bra, rank - rr, np_ptr(i64(L.dims[rr:])), np_ptr(i64(srow[rr:])),
                                          np_ptr(i64(scol[rr:])), L.d_bra_idx.data_ptr(), self.CR.data_ptr(),
                                          self.CC.data_ptr(), st))
        tiles = []
        for s in (range(L.nsec) if sectors is None else sectors):
            r, c = int(L.rows[s]), int(L.cols[s])
            nr, nc = -(-r // RELAYOUT_TR), -(-c // RELAYOUT_TC)
            if nr * nc == 0:
                continue
            r0 = (np.arange(nr, dtype=np.int64) * RELAYOUT_TR)[:, None]
            c0 = (np.arange(nc, dtype=np.int64) * RELAYOUT_TC)[None, :]
            t = np.zeros((nr, nc), dtype=_lib.RELAYOUT_TILE_DTYPE)
            off, ld = int(packed.off[s]), int(packed.ld[s])
            t["dst"] = off + r0 * ld + c0
            t["rbase"] = int(L.ket_start[s]) + r0
            t["cbase"] = int(L.bra_start[s]) + c0
            t["nr"] = np.minimum(RELAYOUT_TR, r - r0)
            t["nc"] = np.minimum(RELAYOUT_TC, c - c0)
            t["dld"] = ld
            tiles.append(t.reshape(-1))
        tiles = np.concatenate(tiles) if tiles else np.zeros(0, dtype=_lib.RELAYOUT_TILE_DTYPE)
        self.ntiles = len(tiles)
        self.d_tiles = torch.from_numpy(tiles.view(np.uint8)).to(dev) if self.ntiles else \
            torch.zeros(32, dtype=torch.uint8, device=dev)
        self.d_index = None
        self.d_runs = None
        self.d_seg, self.nseg = None, 0
        self._seg_args = {}
        self.index_n = 0
        # replay form: element index (fastest: 16.4 us cold / 8.4 us L2-warm on the C3a operand), runs of the
        # index (18.4 / 12.4 us, ~1 % of the index's memory: taken when the index exceeds its budget, or with
        # TOR10_RELAYOUT_RUNS=1), factored tables (24.6 / 16.6 us, no plan memory; TOR10_RELAYOUT_INDEX=0)
        self.use_index = os.environ.get("TOR10_RELAYOUT_INDEX", "1") == "1"
        self.index_fits = 4 * packed.size <= self.INDEX_BUDGET_BYTES
        self.moved_elems = int(sum(int(L.rows[s]) * int(L.cols[s])
                                   for s in (range(L.nsec) if sectors is None else sectors)))

    INDEX_BUDGET_BYTES = int(os.environ.get("TOR10_RELAYOUT_INDEX_MB", "1024")) << 20

    def plan_bytes(self):
        return _tensor_bytes(self.RR, self.RC, self.CR, self.CC, self.d_tiles, self.d_index, self.d_runs, self.d_seg,
                             getattr(self, "d_chunk_first", None))

    def _build_index(self):
        """Element index of the relayout (plan time): the block kernel in EMIT mode."""
        n = (self.packed.size + 1) // 2 * 2
        self.index_n = n
        self.d_index = torch.full((n,), -2, device=self.L.device, dtype=torch.int32)
        check(_lib.load().t10_relayout_blocks(
            _lib.T10_F64, None, None, self.ntiles, self.d_tiles.data_ptr(), self.RR.data_ptr(), self.RC.data_ptr(),
            self.CR.data_ptr(), self.CC.data_ptr(), self.M.d_rowbase.data_ptr(), self.M.d_coloff.data_ptr(),
            self.M.d_ket_map.data_ptr() if self.check else None, self.M.d_bra_map.data_ptr() if self.check else None,
            self.d_index.data_ptr(), _lib.stream_ptr()))
        self._build_segments()
        if self.d_seg is None:
            self._build_runs()

    RUN_CHUNK = 1024      # RR_CHUNK of relayout.cu
    SEG_MAX = 256         # T10_RELAYOUT_SEG_MAX

    def _build_segments(self):
        """Segment form of the element index (plan time, torch ops on the device, 32 M elements at a time):
        maximal pieces contiguous on both sides, cut at SEG_MAX elements, as int32 {dst, src, len, source}.
        Taken when the pieces average at least 16 elements (TOR10_RELAYOUT_MODE=segments: always); the element
        index is then dropped."""
        self.d_seg, self.nseg = None, 0
        mode = os.environ.get("TOR10_RELAYOUT_MODE", "auto")
        if mode not in ("auto", "segments") or os.environ.get("TOR10_RELAYOUT_RUNS", "auto") == "1":
            return
        n = self.index_n
        dev = self.d_index.device
        parts = []
        CH = 1 << 25
        for c0 in range(0, n, CH):
            idx = self.d_index[c0:min(n, c0 + CH)].to(torch.int64)
            m = idx.numel()
            pos = torch.arange(c0, c0 + m, device=dev, dtype=torch.int64)
            key = torch.where(idx >= 0, idx - pos, idx - (1 << 40))
            brk = torch.ones(m, dtype=torch.bool, device=dev)
            brk[1:] = key[1:] != key[:-1]
            st = torch.nonzero(brk).flatten()
            ln = torch.diff(st, append=torch.tensor([m], device=dev, dtype=torch.int64))
            src = idx[st]
            keep = src != -2                                  # untouched destination (padding, foreign sectors)
            st, ln, src = st[keep] + c0, ln[keep], src[keep]
            npc = (ln + self.SEG_MAX - 1) // self.SEG_MAX     # pieces per run
            rep = torch.repeat_interleave(torch.arange(st.numel(), device=dev), npc)
            first = torch.cumsum(npc, 0) - npc
            k = torch.arange(rep.numel(), device=dev) - first[rep]
            off = k * self.SEG_MAX
            seg = torch.empty((rep.numel(), 4), dtype=torch.int32, device=dev)
            seg[:, 0] = (st[rep] + off).to(torch.int32)
            seg[:, 1] = torch.where(src[rep] >= 0, src[rep] + off, src[rep]).to(torch.int32)
            seg[:, 2] = torch.minimum(ln[rep] - off, torch.tensor(self.SEG_MAX, device=dev)).to(torch.int32)
            seg[:, 3] = 0
            parts.append(seg)
        seg = torch.cat(parts) if len(parts) > 1 else parts[0]
        nseg = int(seg.shape[0])
        if nseg == 0 or (mode == "auto" and self.moved_elems < 16 * nseg):
            return
        self.d_seg, self.nseg = seg.contiguous(), nseg
        self.d_index = None                    # 4 B per element of plan memory given back
        self.use_index = False

    def _build_runs(self):
        """Run-length form of the element index (plan time, torch ops on the device): maximal pieces with
        index[i+1] == index[i] + 1 (or constant -1 / -2).  Used when runs average >= 16 elements; the element
        index is then dropped."""
        self.d_runs = None
        want = os.environ.get("TOR10_RELAYOUT_RUNS", "auto")
        if want == "0" or (want == "auto" and self.index_fits):
            if not self.index_fits:
                self.d_index, self.use_index = None, False     # over budget and no runs: factored tables
            return
        n = self.index_n
        idx = self.d_index.to(torch.int64)
        pos = torch.arange(n, device=idx.device, dtype=torch.int64)
        key = torch.where(idx >= 0, idx - pos, idx - (1 << 40))
        brk = torch.ones(n, dtype=torch.bool, device=idx.device)
        brk[1:] = key[1:] != key[:-1]
        starts = torch.nonzero(brk).flatten()
        nruns = int(starts.numel())
        if nruns * 16 > n:
            if not self.index_fits:
                self.d_index, self.use_index = None, False
            return
        runs = torch.empty((nruns + 1, 2), dtype=torch.int32, device=idx.device)
        runs[:nruns, 0] = starts.to(torch.int32)
        runs[:nruns, 1] = idx[starts].to(torch.int32)
        runs[nruns, 0], runs[nruns, 1] = n, 0
        nchunks = (n + self.RUN_CHUNK - 1) // self.RUN_CHUNK
        first = torch.clamp(torch.arange(nchunks + 1, device=idx.device, dtype=torch.int64) * self.RUN_CHUNK, max=n - 1)
        self.d_chunk_first = (torch.searchsorted(starts, first, right=True) - 1).to(torch.int32)
        self.d_runs, self.nruns = runs.contiguous(), nruns
        self.d_index = None                    # 4 B per element of plan memory given back
        self.use_index = False

    def run_segments(self, src_ptrs, dst, dcode, seg=None, nseg=None):
        """Segment replay; src_ptrs[k] = base pointer of source arena k (one entry on a single GPU)."""
        key = tuple(src_ptrs)
        arr = self._seg_args.get(key)
        if arr is None:
            if len(self._seg_args) > 32:
                self._seg_args.clear()
            arr = self._seg_args[key] = (C.c_void_p * len(src_ptrs))(*src_ptrs)
        seg = self.d_seg if seg is None else seg
        check(_lib.load().t10_relayout_segments(dcode, len(src_ptrs), arr, dst.data_ptr(),
                                                self.nseg if nseg is None else nseg, seg.data_ptr(),
                                                _lib.stream_ptr()))

    def run(self, src_arena, dst):
        """dst must hold packed.size elements; padding/unmatched space is left untouched."""
        if self.use_index and self.d_index is None and self.d_runs is None and self.d_seg is None and self.ntiles \
                and src_arena.numel() < 2 ** 31:
            # an index over its memory budget is only built (transiently) when its run-length form was asked for;
            # otherwise such a plan replays through the factored tables, which need no plan memory at all
            if self.index_fits or os.environ.get("TOR10_RELAYOUT_RUNS", "auto") == "1":
                with torch.cuda.device(self.L.device):
                    self._build_index()
                _RELAYOUTS.recount()
            else:
                self.use_index = False
        if self.d_seg is not None and dst.numel() >= self.index_n:
            self.run_segments([src_arena.data_ptr()], dst, _lib.dtype_code(src_arena.dtype))
            return dst
        if self.d_runs is not None and dst.numel() >= self.index_n and dst.data_ptr() % 16 == 0:
            check(_lib.load().t10_relayout_runs(_lib.dtype_code(src_arena.dtype), src_arena.data_ptr(), dst.data_ptr(),
                                                self.index_n, self.d_runs.data_ptr(), self.nruns,
                                                self.d_chunk_first.data_ptr(), _lib.stream_ptr()))
            return dst
        if self.d_index is not None and dst.numel() >= self.index_n and dst.data_ptr() % 16 == 0:
            check(_lib.load().t10_relayout_index(_lib.dtype_code(src_arena.dtype), src_arena.data_ptr(),
                                                 dst.data_ptr(), self.index_n, self.d_index.data_ptr(),
                                                 _lib.stream_ptr()))
            return dst
        check(_lib.load().t10_relayout_blocks(
            _lib.dtype_code(src_arena.dtype), src_arena.data_ptr(), dst.data_ptr(), self.ntiles,
            self.d_tiles.data_ptr(), self.RR.data_ptr(), self.RC.data_ptr(),
            self.CR.data_ptr(), self.CC.data_ptr(), self.M.d_rowbase.data_ptr(), self.M.d_coloff.data_ptr(),
            self.M.d_ket_map.data_ptr() if self.check else None, self.M.d_bra_map.data_ptr() if self.check else None,
            None, _lib.stream_ptr()))
        return dst


def get_relayout(M, L, mapper, padded=False, sectors=None):
    key = (M.key, L.key, tuple(int(x) for x in mapper), bool(padded), None if sectors is None else tuple(sectors))
    pl = _RELAYOUTS.get(key)
    if pl is None:
        packed = PackedLayout.padded(L, sectors) if padded else PackedLayout.native(L)
        pl = RelayoutPlan(M, L, mapper, packed, sectors)
        _RELAYOUTS[key] = pl
    return pl


# ------------------------------------------------------------------------------------
def label_match(la, lb):
    """np.intersect1d(..., return_indices=True) + setdiff1d (UniTensor.py:3686-3689): common labels
    in ASCENDING label order (quirk Q5), their positions, and the free positions ascending."""
    same, a_ind, b_ind = np.intersect1d(la, lb, return_indices=True)
    a_free = np.setdiff1d(np.arange(len(la)), a_ind)
    b_free = np.setdiff1d(np.arange(len(lb)), b_ind)
    return same, a_ind.astype(np.int64), b_ind.astype(np.int64), a_free.astype(np.int64), b_free.astype(np.int64)


def pick_tile_cfg(problems, sharded=False, base_aligned=False):
    """Tile configuration of the FP64 grouped GEMM (see t10_ggemm_plan_tiles), from measurements on B200
    (profiles/r01_lone_cta.jsonl, r01_bench_*.json):
      3   64x64x16, 3 stages, 3 CTAs/SM, any alignment (8-byte cp.async on odd leading dimensions), ragged
          fragments skipped, static LPT slots -- the general kernel (C3a: 129.8 us);
      40  same tile on the software-pipelined aligned-only kernel, 2 CTAs/SM: 33 TF against 30.6 TF in steady
          state, C3a GEMM stage 123.9 us -- taken when every operand row is 16-byte aligned AS STORED (an
          extra copy of an operand just to align it costs more than it gains);
      42  64x128 tiles on the same kernel (33.9 TF steady state) for aligned groups of at least four waves of
          such tiles (large dense GEMMs, the HBM-bound small-K sectors of C5);
      46  the same main loop with mixed tile shapes (64x64 + 32x128 / 128x32 edge strips) for aligned GROUPS
          (block-sparse contractions), run as stream-K when every CTA gets at least 16 k-steps (C3a GEMM stage
          114.7 -> 110.9 us);
      8   32x32 tiles for a rank of a sector-sharded contraction left with fewer than 1.5 64x64 tiles per SM
          (C3a on 8 GPUs: 53.5 us -> 39.6 us)."""
    forced = os.environ.get("TOR10_GGEMM_CFG")
    if forced is not None:
        return int(forced)
    if problems is None or not len(problems):
        return 3
    m, n = problems["m"].astype(np.int64), problems["n"].astype(np.int64)
    aligned = base_aligned and bool(((problems["lda"] % 2 == 0) & (problems["ldb"] % 2 == 0)
                                     & (problems["a_off"] % 2 == 0) & (problems["b_off"] % 2 == 0)).all())
    few = sharded and int((((m + 63) // 64) * ((n + 63) // 64)).sum()) < 1.5 * _sm_count()
    if few and not (aligned and os.environ.get("TOR10_SHARDED_SK", "1") == "1"):
        return 8
    if not aligned:
        return 3
    if sharded:
        # a rank's share of the sectors: stream-K on the mixed-shape kernel when every CTA still gets 16 k-steps
        # (its all-gather staging tile aliases the pipeline stages, so two CTAs stay resident), else plain 64x64
        ksteps = int((((m + 63) // 64) * ((n + 63) // 64) * ((problems["k"].astype(np.int64) + 15) // 16)).sum())
        if os.environ.get("TOR10_SHARDED_SK", "1") == "1" and ksteps >= 16 * 2 * _sm_count():
            return 46
        return 8 if few else 40
    if int((((m + 63) // 64) * ((n + 127) // 128)).sum()) >= 8 * _sm_count():
        return 42      # many waves of work: the larger tile wins (C5, K <= 64 per sector: 30.7 ms against 38.2 ms)
    if len(problems) > 1:
        return 46
    return 40


def _sm_count_of(device):
    with torch.cuda.device(device):
        return _sm_count()


def _sm_count():
    nsm, z = C.c_int(0), C.c_int(0)
    check(_lib.load().t10_device_props(torch.cuda.current_device(), C.byref(nsm), C.byref(z), C.byref(z), C.byref(z)))
    return int(nsm.value)


ALIGNED_CFGS = (40, 42, 46)   # kernels that require 16-byte aligned operand rows
FUSABLE_CFGS = (0, 2, 3, 8, 40, 42, 46)   # kernels with the fused all-gather epilogue
FALLBACK_CFG = 3


class GemmGroup:
    """Problem table + tile table on the device."""

    def __init__(self, problems, device, dtype, tile_cfg=None, base_aligned=True, sharded=False, resident=False,
                 force_tma=False):
        lib = _lib.load()
        if resident:
            sharded = False       # a rank's share of a resident contraction is an ordinary group: no fused push
        self.problems = np.ascontiguousarray(problems, dtype=_lib.GEMM_PROBLEM_DTYPE)
        self.nprob = len(self.problems)
        self.dtype_code = _lib.dtype_code(dtype)
        auto_cfg = tile_cfg is None and os.environ.get("TOR10_GGEMM_CFG") is None
        if tile_cfg is None:
            if dtype == torch.float64:
                with torch.cuda.device(device):
                    tile_cfg = pick_tile_cfg(self.problems, sharded=sharded, base_aligned=bool(base_aligned))
            else:   # fp32: tcgen05 3xTF32 kernel (50, default) or the SIMT kernel (0)
                tile_cfg = int(os.environ.get("TOR10_F32_CFG", "50"))
        self.tile_cfg = int(tile_cfg)
        pr = self.problems
        self.aligned = bool(((pr["lda"] % 2 == 0) & (pr["ldb"] % 2 == 0) & (pr["a_off"] % 2 == 0)
                             & (pr["b_off"] % 2 == 0)).all()) and bool(base_aligned)
        if self.tile_cfg in ALIGNED_CFGS and not self.aligned:
            self.tile_cfg = FALLBACK_CFG      # odd leading dimension somewhere: predicated 8-byte loader kernel
        # CTA slots of the persistent kernel (occupancy query: needs the device) -> LPT slot table
        ns = C.c_int(0)
        with torch.cuda.device(device):
            check(lib.t10_ggemm_slots(self.dtype_code, self.tile_cfg, C.byref(ns)))
        # schedule: "smbins" = one LPT bin per SM shared by the SM's resident CTAs through counters (kernels
        # 0..3), "slots" = one static LPT slot per resident CTA, "dynamic" = one global counter
        mode = os.environ.get("TOR10_GGEMM_SCHED", "slots")
        f64k = dtype == torch.float64 and self.tile_cfg != 100
        counters_ok = f64k and self.tile_cfg not in ALIGNED_CFGS
        smbins = mode == "smbins" and counters_ok
        use_slots = f64k and (mode == "slots" or smbins or (mode in ("dynamic", "smbins") and not counters_ok))
        nt = C.c_int64(0)
        check(lib.t10_ggemm_plan_tiles(self.tile_cfg, self.nprob, np_ptr(self.problems), 0, 0, None, None,
                                       C.byref(nt)))
        if smbins:
            nsm = C.c_int(0)
            z = C.c_int(0)
            check(lib.t10_device_props(device.index if device.index is not None else torch.cuda.current_device(),
                                       C.byref(nsm), C.byref(z), C.byref(z), C.byref(z)))
            self.nslots = int(nsm.value)
            self.d_sched = torch.zeros(2 + self.nslots, dtype=torch.int32, device=device)
        else:
            self.nslots = min(int(ns.value), max(int(nt.value), 1)) if use_slots else 0
            self.d_sched = torch.zeros(2, dtype=torch.int32, device=device) \
                if (mode == "dynamic" and counters_ok) else None
        tiles = np.zeros((max(nt.value, 1), 4), dtype=np.int32)
        slots = np.zeros(self.nslots + 1, dtype=np.int32)
        check(lib.t10_ggemm_plan_tiles(self.tile_cfg, self.nprob, np_ptr(self.problems), self.nslots, nt.value,
                                       np_ptr(tiles), np_ptr(slots) if self.nslots else None, C.byref(nt)))
        self.ntiles = int(nt.value)
        self.d_tiles = torch.from_numpy(tiles).to(device)
        self.d_slots = torch.from_numpy(slots).to(device) if self.nslots else None
        self.d_prob = torch.from_numpy(self.problems.view(np.uint8).reshape(-1)).to(device) if self.nprob else \
            torch.zeros(48, dtype=torch.uint8, device=device)
        self.streamk = False
        self.tma = False
        self._tma_maps = {}
        if (dtype == torch.float64 and self.tile_cfg in (46, 40) and self.ntiles > 0 and self.aligned
                and not sharded and os.environ.get("TOR10_GGEMM_TMA", "1") == "1"
                # a lone small GEMM (the chi = 32 products of iTEBD) gains nothing from the TMA feed and would pay
                # for a tensor-map encode + upload whenever its operands move: the cp.async kernel takes those
                # (force_tma: the "push" transport needs the same kernel on every rank, whatever its share)
                and (force_tma or self.nprob > 1 or self.ntiles >= 2 * _sm_count_of(device))):
            # TMA-fed kernel (ggemm_tma.cu): stream-K pieces when there is enough work to cut, whole tiles in LPT
            # slots otherwise -- one kernel, one table format
            nres = C.c_int(0)
            with torch.cuda.device(device):
                check(lib.t10_ggemm_tma_slots(C.byref(nres)))
            self.tma = self.build_streamk(int(nres.value), phases=1, min_ksteps=int(os.environ.get("TOR10_TMA_MINK", "4")),
                                          fixed_default=float(os.environ.get("TOR10_TMA_FIXED", "1"))) \
                or self.build_whole_tiles(int(nres.value))
        if (not self.tma and dtype == torch.float64 and self.tile_cfg == 46 and self.ntiles > 0
                and os.environ.get("TOR10_GGEMM_STREAMK", "1") in ("1", "force")):
            # a sharded rank pushes its tiles to the peers as they finish: two groups of tiles, so that half of them
            # finish (and travel) while the other half computes
            self.streamk = self.build_streamk(int(ns.value), phases=int(os.environ.get("TOR10_SK_PHASES", "1"))
                                              if sharded else 1)
        if sharded and auto_cfg and self.tile_cfg == 46 and not self.streamk:
            # too little work to cut tiles after all: the plain 64x64 (or 32x32) kernels
            few = int((((pr["m"].astype(np.int64) + 63) // 64) * ((pr["n"].astype(np.int64) + 63) // 64)).sum()) \
                < 1.5 * _sm_count()
            return self.__init__(problems, device, dtype, tile_cfg=8 if few else 40, base_aligned=base_aligned,
                                 sharded=sharded)
        self.flops = float(sum(2.0 * int(p["m"]) * int(p["n"]) * int(p["k"]) for p in self.problems))
        self.bytes = float(sum(int(p["m"]) * int(p["k"]) + int(p["k"]) * int(p["n"]) + int(p["m"]) * int(p["n"])
                               for p in self.problems))

    # ---- stream-K on the mixed-shape kernel: balance k-steps, not tiles -------------------------------------
    def build_whole_tiles(self, nslots):
        """Whole tiles (mixed shapes) dealt to CTA slots by LPT, in the piece-table format of the stream-K / TMA
        kernels: every piece is a complete tile (no parked partials)."""
        lib = _lib.load()
        nt = C.c_int64(0)
        check(lib.t10_ggemm_plan_tiles(46, self.nprob, np_ptr(self.problems), 0, 0, None, None, C.byref(nt)))
        if nt.value == 0:
            return False
        nslots = max(1, min(int(nslots), int(nt.value)))
        tl = np.zeros((nt.value, 4), dtype=np.int32)
        slots = np.zeros(nslots + 1, dtype=np.int32)
        check(lib.t10_ggemm_plan_tiles(46, self.nprob, np_ptr(self.problems), nslots, nt.value, np_ptr(tl),
                                       np_ptr(slots), C.byref(nt)))
        kts = (self.problems["k"][tl[:, 0]].astype(np.int64) + 15) // 16
        pcs = np.zeros((len(tl), 4), dtype=np.int32)
        pcs[:, 1] = kts
        pcs[:, 2] = -1
        dev = self.d_prob.device
        self.sk_phases, self.sk_parked = 1, 0
        self.sk_nslots, self.sk_npieces = nslots, len(tl)
        self.d_sk_tiles = torch.from_numpy(tl).to(dev)
        self.d_sk_pieces = torch.from_numpy(pcs).to(dev)
        self.d_sk_slots = torch.from_numpy(slots).to(dev)
        self.d_sk_ws = torch.empty(128 * 34, dtype=torch.float64, device=dev)
        self.d_sk_flags = torch.zeros(nslots + 1, dtype=torch.int32, device=dev)
        self.sk_epoch = 0
        self._sk_static = None
        return True

    def build_streamk(self, nslots, phases=1, min_ksteps=16, fixed_default=2.0):
        """Cut the tile list (t10_ggemm_plan_tiles, cfg 46, heaviest first) into `nslots` equal shares of k-steps
        (a k-step of the aligned kernel costs the same whatever the tile's raggedness; +2 per piece for the
        pipeline fill).  A tile cut across CTAs becomes pieces: the CTAs holding its leading k-ranges park partial
        sums, the one holding its last k-step adds them in order and stores the tile (t10_ggemm_streamk).  Per
        CTA the parked piece goes first and the piece that waits for parked partials goes last."""
        lib = _lib.load()
        nt = C.c_int64(0)
        check(lib.t10_ggemm_plan_tiles(46, self.nprob, np_ptr(self.problems), 0, 0, None, None, C.byref(nt)))
        if nt.value == 0:
            return False
        tl = np.zeros((nt.value, 4), dtype=np.int32)
        check(lib.t10_ggemm_plan_tiles(46, self.nprob, np_ptr(self.problems), 0, nt.value, np_ptr(tl), None, C.byref(nt)))
        fixed = float(os.environ.get("TOR10_SK_FIXED", str(fixed_default)))   # measured on C3a: (2, 2) 109.9 us, (3, 2) 110.9,
        min_piece = int(os.environ.get("TOR10_SK_MINPIECE", "2"))     # (2, 1) 109.7, (2, 4) 111.5, (5, 2) 111.9, (2, 8) 120.1
        if os.environ.get("TOR10_SK_ORDER", "cost") == "spread":
            # low-discrepancy interleaving of the cost-sorted list: every CTA's share then mixes long and short
            # tiles (the short ones carry more pipeline-fill overhead than the model charges)
            key = (np.arange(len(tl)) * 0.6180339887498949) % 1.0
            tl = tl[np.argsort(key, kind="stable")]
        kts = (self.problems["k"][tl[:, 0]].astype(np.int64) + 15) // 16
        if int(kts.sum()) < min_ksteps * nslots and os.environ.get("TOR10_GGEMM_STREAMK") != "force":
            return False          # too little work per CTA to be worth cutting tiles (plain LPT slots instead)
        # PHASES (sharded ranks with the fused all-gather): the tiles are dealt into `phases` groups of equal work and
        # every group is cut into nslots shares of its own, run one group after the other.  With one group all CTAs
        # finish their last tile at the same moment and most of the result is pushed to the peers in the kernel's
        # last microseconds; with G groups (G - 1) / G of it is pushed while the next group computes.
        phases = max(1, int(phases))
        while phases > 1 and int(kts.sum()) < 8 * phases * nslots:
            phases -= 1
        tiles, pcs = [[] for _ in range(nslots)], [[] for _ in range(nslots)]
        nparked = 0
        for g in range(phases):
            sel = np.arange(g, len(tl), phases)          # every phases-th tile of the cost-sorted list
            w = kts[sel] + fixed
            share = float(w.sum()) / nslots
            per_cta = [[] for _ in range(nslots)]      # entries (tile, kt0, kt1)
            pos = 0.0
            for j, ti in enumerate(sel):
                kt = int(kts[ti])
                lo, hi = pos, pos + float(w[j])
                pos = hi
                c0 = min(nslots - 1, int(lo / share + 1e-9))
                c1 = min(nslots - 1, int((hi - 1e-9) / share))
                if c0 == c1 or kt < 2 * min_piece:
                    per_cta[min(nslots - 1, int((0.5 * (lo + hi)) / share))].append((ti, 0, kt))
                    continue
                cuts = [0]
                for c in range(c0 + 1, c1 + 1):        # k-step at which CTA c takes over
                    cuts.append(max(cuts[-1], min(kt, int(round(c * share - lo - fixed)))))
                cuts.append(kt)
                merged = []
                for c, a_, b_ in ((c0 + i, cuts[i], cuts[i + 1]) for i in range(c1 - c0 + 1)):
                    if b_ - a_ < min_piece and merged:
                        merged[-1] = (merged[-1][0], merged[-1][1], b_)
                    else:
                        if merged and merged[-1][2] - merged[-1][1] < min_piece:
                            a_ = merged.pop()[1]
                        merged.append((c, a_, b_))
                if len(merged) > 255:
                    return False
                for c, a_, b_ in merged:
                    per_cta[c].append((ti, a_, b_))
            holders = {}
            for c in range(nslots):
                for ti, a_, b_ in per_cta[c]:
                    holders.setdefault(ti, []).append(c)
            slot0 = g * (nslots + 1)                   # park slots of this phase (flag nslots reports errors)
            for c in range(nslots):
                parked = [x for x in per_cta[c] if x[2] < int(kts[x[0]])]
                rest = [x for x in per_cta[c] if x[2] >= int(kts[x[0]])]
                if len(parked) > 1:
                    return False                       # cannot happen with contiguous shares
                nparked += len(parked)
                for ti, a_, b_ in parked + [x for x in rest if x[1] == 0] + [x for x in rest if x[1] > 0]:
                    tiles[c].append(tl[ti])
                    if b_ < int(kts[ti]):
                        pcs[c].append((a_, b_, slot0 + c, 0))
                    elif a_ > 0:
                        hs = holders[ti]
                        nfix = len(hs) - 1
                        if hs != list(range(hs[0], hs[0] + nfix + 1)) or hs[-1] != c:
                            return False
                        pcs[c].append((a_, b_, -1, nfix | ((slot0 + hs[0]) << 8)))
                    else:
                        pcs[c].append((0, b_, -1, 0))
        starts = [0]
        for c in range(nslots):
            starts.append(starts[-1] + len(tiles[c]))
        tiles = [x for c in range(nslots) for x in tiles[c]]
        pcs = [x for c in range(nslots) for x in pcs[c]]
        self.sk_phases = phases
        dev = self.d_prob.device
        self.sk_nslots, self.sk_npieces = nslots, len(tiles)
        self.d_sk_tiles = torch.from_numpy(np.ascontiguousarray(np.array(tiles, dtype=np.int32).reshape(-1, 4))).to(dev)
        self.d_sk_pieces = torch.from_numpy(np.array(pcs, dtype=np.int32).reshape(-1, 4)).to(dev)
        self.d_sk_slots = torch.from_numpy(np.array(starts, dtype=np.int32)).to(dev)
        self.d_sk_ws = torch.empty(phases * (nslots + 1) * 128 * 34, dtype=torch.float64, device=dev)
        self.d_sk_flags = torch.zeros(phases * (nslots + 1), dtype=torch.int32, device=dev)
        self.sk_epoch = 0
        self._sk_static = None
        self.sk_parked = nparked
        return True

    def plan_bytes(self):
        """Tile / piece tables, the stream-K workspace of parked partials (~10 MB), cached tensor maps."""
        return _owned_tensor_bytes(self)

    def run_streamk(self, A, B, Cc):
        self.sk_epoch = self.sk_epoch % 2000000000 + 1
        st = self._sk_static
        if st is None:       # pointers of the plan's own tables never change: resolve them once
            st = self._sk_static = (_lib.load().t10_ggemm_streamk, self.d_prob.data_ptr(), self.d_sk_tiles.data_ptr(),
                                    self.d_sk_pieces.data_ptr(), self.d_sk_slots.data_ptr(), self.d_sk_ws.data_ptr(),
                                    self.d_sk_flags.data_ptr())
        check(st[0](A.data_ptr(), B.data_ptr(), Cc.data_ptr(), self.nprob, st[1], self.sk_npieces, st[2], st[3],
                    self.sk_nslots, st[4], st[5], st[6], self.sk_epoch, _lib.stream_ptr()))

    def run_tma(self, A, B, Cc, pdl=False, signal=None):
        """TMA-fed kernel: the tensor maps (one per sector and operand) are encoded on the host for the operands'
        base addresses and cached -- a loop that contracts the same tensors, or whose scratch buffers come back from
        the caching allocator at the same address, pays for them once."""
        lib = _lib.load()
        key = (A.data_ptr(), B.data_ptr())
        d_maps = self._tma_maps.get(key)
        if d_maps is None:
            h = torch.empty(max(self.nprob, 1) * 256, dtype=torch.uint8,
                            pin_memory=(A.device.type == "cuda"))
            check(lib.t10_ggemm_tma_encode(self.nprob, np_ptr(self.problems), A.data_ptr(), B.data_ptr(), h.data_ptr()))
            d_maps = h.to(A.device, non_blocking=True)
            if len(self._tma_maps) >= 16:
                self._tma_maps.pop(next(iter(self._tma_maps)))
            self._tma_maps[key] = d_maps
            if A.is_cuda and torch.cuda.is_current_stream_capturing():
                # a captured upload re-reads its pinned source on every replay: keep it alive with the graph's maps
                self._tma_pinned = getattr(self, "_tma_pinned", []) + [h]
        if A.is_cuda and torch.cuda.is_current_stream_capturing():
            # the graph addresses these maps by raw pointer: they must outlive their slot in the 16-entry cache
            pinned = self.__dict__.setdefault("_tma_pinned", [])
            if not any(p is d_maps for p in pinned):
                pinned.append(d_maps)
        if self.sk_parked and A.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("tor10_b200: this contraction cuts tiles across CTAs (stream-K flags carry a per-launch "
                               "epoch) and cannot be captured into a CUDA graph; run it eagerly")
        self.sk_epoch = self.sk_epoch % 2000000000 + 1
        nsig, sig_ptrs, sig_val, done = (0, None, 0, None) if signal is None else \
            (len(signal[0]), signal[0], int(signal[1]), signal[2].data_ptr())
        push = signal[3] if (signal is not None and len(signal) > 3) else None
        check(lib.t10_ggemm_tma(d_maps.data_ptr(), Cc.data_ptr(), self.nprob, self.d_prob.data_ptr(), self.sk_npieces,
                                self.d_sk_tiles.data_ptr(), self.d_sk_pieces.data_ptr(), self.sk_nslots,
                                self.d_sk_slots.data_ptr(), self.d_sk_ws.data_ptr(), self.d_sk_flags.data_ptr(),
                                self.sk_epoch, 1 if pdl else 0, nsig, sig_ptrs, sig_val, done,
                                len(push) if push is not None else 0, push, _lib.stream_ptr()))

    def fusable(self):
        """The fused GEMM -> all-gather epilogue exists in the fp64 kernels 0..3 (t10_ggemm_allgather)."""
        return self.dtype_code == _lib.dtype_code(torch.float64) and self.tile_cfg in FUSABLE_CFGS

    def run_allgather(self, A, B, peer_ptrs, flag_ptrs, step, done_ctr, self_index=-1):
        """peer_ptrs / flag_ptrs: ctypes (c_void_p * n) of every rank's arena base / of this rank's flag slot
        on every rank (this rank's own included, at position self_index)."""
        if self.streamk:
            self.sk_epoch = self.sk_epoch % 2000000000 + 1
            check(_lib.load().t10_ggemm_streamk_allgather(
                A.data_ptr(), B.data_ptr(), len(peer_ptrs), self_index, peer_ptrs, flag_ptrs, step, done_ctr.data_ptr(),
                self.nprob, self.d_prob.data_ptr(), self.sk_npieces, self.d_sk_tiles.data_ptr(),
                self.d_sk_pieces.data_ptr(), self.sk_nslots, self.d_sk_slots.data_ptr(), self.d_sk_ws.data_ptr(),
                self.d_sk_flags.data_ptr(), self.sk_epoch, _lib.stream_ptr()))
            return
        check(_lib.load().t10_ggemm_allgather(self.tile_cfg, A.data_ptr(), B.data_ptr(), len(peer_ptrs), self_index,
                                              peer_ptrs, flag_ptrs, step, done_ctr.data_ptr(),
                                              self.nprob, self.d_prob.data_ptr(), self.ntiles,
                                              self.d_tiles.data_ptr(), self.nslots,
                                              self.d_slots.data_ptr() if self.nslots else None,
                                              self.d_sched.data_ptr() if self.d_sched is not None else None,
                                              _lib.stream_ptr()))

    def run(self, A, B, Cc, pdl=False, signal=None):
        """signal = (ctypes array of progress-word addresses, value, arrival counter): published when the launch
        has stored its last tile (sector-resident multi-GPU runs)."""
        if self.tma:
            return self.run_tma(A, B, Cc, pdl=pdl, signal=signal)
        if signal is not None:
            self.run(A, B, Cc)
            check(_lib.load().t10_signal_flags(len(signal[0]), signal[0], int(signal[1]), _lib.stream_ptr()))
            return
        if self.streamk:
            return self.run_streamk(A, B, Cc)
        check(_lib.load().t10_ggemm(self.dtype_code, self.tile_cfg, A.data_ptr(), B.data_ptr(), Cc.data_ptr(),
                                    self.nprob, self.d_prob.data_ptr(), self.ntiles, self.d_tiles.data_ptr(),
                                    self.nslots, self.d_slots.data_ptr() if self.nslots else None,
                                    self.d_sched.data_ptr() if self.d_sched is not None else None,
                                    _lib.stream_ptr()))


_PLAN_SERIAL = itertools.count()
_PDL = os.environ.get("TOR10_PDL", "1") == "1"
_SHARD_MIN_FLOP = float(os.environ.get("TOR10_SHARD_MIN_MFLOP", "5")) * 1e6


class ContractPlan:
    """Symmetric Contract(a, b): UniTensor.py:3692-3743."""

    def __init__(self, a, b, shard=None):
        self.serial = next(_PLAN_SERIAL)     # identity of the plan's private staging arenas (never reused)
        from .Bond import BondType, BD_KET
        STATS["contract_build"] += 1
        self.shard = shard
        self._transport, self._stage = None, None
        if shard is not None:
            from . import parallel
            self._transport = parallel.transport_for()
        same, a_ind, b_ind, a_free, b_free = label_match(a.labels, b.labels)
        if len(same) == 0:
            raise Exception("Developing")                                          # :3745-3747 (Q6)
        if ((a.braket[a_ind] + b.braket[b_ind]) != 0).any():                        # :3700-3701
            raise Exception("Contract(a,b)", "[ERROR] bra-bond can only contract with ket-bond")
        dev = a.device
        perm_a = np.append(a_free, a_ind)
        perm_b = np.append(b_ind, b_free)
        rr_a, rr_b = len(a_free), len(b_ind)
        if rr_a < 1 or len(b_free) < 1:
            # the reference builds `out` from zero row (col) bonds and fails in UniTensor.__init__
            raise TypeError("UniTensor.__init__", "[ERROR] tensor with symmetry must have at least one rank for "
                                                  "row space and one rank for column space")
        bonds_a = [a.bonds[i] for i in perm_a]
        bonds_b = [b.bonds[i] for i in perm_b]
        bk_a, bk_b = a.braket[perm_a], b.braket[perm_b]
        map_a, map_b = a._mapper[perm_a], b._mapper[perm_b]
        Ma, Mb = a._layout, b._layout
        self._keep = (Ma, Mb)
        La = get_layout(bonds_a, bk_a, rr_a, dev)
        Lb = get_layout(bonds_b, bk_b, rr_b, dev)
        # output: a's free bonds then b's free bonds; rowrank INFERRED as #KET (quirk Q2, :3720-3723)
        self.out_bonds = bonds_a[:rr_a] + bonds_b[rr_b:]
        self.out_labels = np.append(a.labels[perm_a][:rr_a], b.labels[perm_b][rr_b:])
        self.out_braket = np.append(bk_a[:rr_a], bk_b[rr_b:])
        self.out_rowrank = int((self.out_braket == BondType[BD_KET]).sum())
        if self.out_rowrank < 1 or self.out_rowrank >= len(self.out_bonds):
            raise TypeError("UniTensor.__init__", "[ERROR] tensor with symmetry must have at least one rank for "
                                                  "row space and one rank for column space")
        # The GEMMs produce the result split as (a's free bonds | b's free bonds).  When the inferred rowrank
        # differs (outside the reference's well-defined domain: there it raises inside torch.matmul or stores
        # wrong-shaped blocks, quirk Q2) the result is returned in that memory layout with the reference's
        # metadata (labels, bonds, inferred rowrank); reading it relayouts like any permuted tensor.
        self.mem_rowrank = rr_a
        Lo = get_layout(self.out_bonds, self.out_braket, self.mem_rowrank, dev)
        self.La, self.Lb, self.Lo = La, Lb, Lo
        # the reference tests only `_mapper == identity` (quirk Q1); a changed rowrank with identity
        # mapper leaves its Storage stale.  Here the relayout runs whenever the memory layout differs.
        # (exchanged bra/ket tags -- Whole_transpose -- also change the sector table: compare them too)
        ident_a = (map_a == np.arange(len(map_a))).all() and rr_a == Ma.rowrank and np.array_equal(Ma.braket, bk_a)
        ident_b = (map_b == np.arange(len(map_b))).all() and rr_b == Mb.rowrank and np.array_equal(Mb.braket, bk_b)
        # sector matching (UniTensor.py:3727-3738)
        matched = []
        for ob in range(Lo.nsec):
            ab = La.sector_of(Lo.block_qnums[ob])
            bb = Lb.sector_of(Lo.block_qnums[ob])
            if ab is not None and bb is not None:
                ma, ka, kb, nb = int(La.rows[ab]), int(La.cols[ab]), int(Lb.rows[bb]), int(Lb.cols[bb])
                if ka != kb or ma != int(Lo.rows[ob]) or nb != int(Lo.cols[ob]):
                    # contracted bonds whose charge tables differ (the reference never compares them, quirk Q4)
                    raise RuntimeError("Contract(a,b): sector %s has mismatched block shapes A[%d,%d] B[%d,%d] "
                                       "out[%d,%d] (the contracted bonds of a and b carry different charges)"
                                       % (Lo.block_qnums[ob].tolist(), ma, ka, kb, nb, Lo.rows[ob], Lo.cols[ob]))
                matched.append((ob, ab, bb))
        self.all_matched = len(matched) == Lo.nsec
        self.flops_total = float(sum(2.0 * int(La.rows[ab]) * int(La.cols[ab]) * int(Lb.cols[bb])
                                     for _, ab, bb in matched))
        self.ranges = None
        if shard is not None:
            # independent sectors are partitioned over the ranks: contiguous runs of output
            # sectors with balanced flops, so that every rank's result is ONE slice of the arena
            from .parallel import partition_contiguous, partition_lpt, arena_ranges
            rank, world = shard
            cost = np.zeros(Lo.nsec)
            for ob, ab, bb in matched:
                cost[ob] = 2.0 * int(La.rows[ab]) * int(La.cols[ab]) * int(Lb.cols[bb])
            bounds = partition_contiguous(cost, world)
            self.sector_bounds = bounds
            self.ranges = arena_ranges(Lo, bounds)
            owner = np.searchsorted(np.asarray(bounds[1:]), np.arange(Lo.nsec), side="right")
            self.inherited = False
            if self._transport in ("resident", "pull", "push") and a._res is not None and ident_a \
                    and a._res.world == world:
                # ownership propagates: output sector q lives where A's row sector q lives, so a chain
                # of contractions reads its running operand in place on every rank
                from .parallel import inherit_owner
                owner = inherit_owner(a._res.owner, matched, Lo.nsec, cost, world)
                self.inherited = True
            elif self._transport in ("none", "resident", "pull", "push") or (self._transport == "fused" and a.dtype == torch.float64
                                             and pick_tile_cfg(None) in FUSABLE_CFGS):
                # the fused GEMM -> all-gather stores tiles, not arena slices (and sharded results are not gathered
                # at all): ownership need not be contiguous, so the sectors are balanced by LPT (the contiguous
                # split of C3a at 2 ranks is 44 / 56 %)
                owner = partition_lpt(cost, world)
            self.sector_owner = owner
            self._own_valid = (np.asarray(owner) == rank) | (np.asarray(owner) < 0)
            self.matched_all = list(matched)
            matched = [m for m in matched if owner[m[0]] == rank]
        self.matched = matched
        sec_a = sorted(set(m[1] for m in matched))
        sec_b = sorted(set(m[2] for m in matched))
        self.sec_a, self.sec_b = sec_a, sec_b
        # (stored sectors have an even leading dimension, so an operand read in place is always 16-byte aligned)
        self.rel_a = None if ident_a else get_relayout(Ma, La, map_a, padded=True, sectors=sec_a)
        self.rel_b = None if ident_b else get_relayout(Mb, Lb, map_b, padded=True, sectors=sec_b)
        pa = self.rel_a.packed if self.rel_a else PackedLayout.native(La)
        pb = self.rel_b.packed if self.rel_b else PackedLayout.native(Lb)
        prob = np.zeros(len(matched), dtype=_lib.GEMM_PROBLEM_DTYPE)
        for i, (ob, ab, bb) in enumerate(matched):
            prob[i] = (pa.off[ab], pb.off[bb], Lo.arena_off[ob], La.rows[ab], Lb.cols[bb], La.cols[ab],
                       pa.ld[ab], pb.ld[bb], Lo.ld[ob])
        self.gemm = GemmGroup(prob, dev, a.dtype,
                              # "push": the same kernel on every rank whatever its share (the ranks must agree on who pushes)
                              tile_cfg=46 if (self._transport == "push" and a.dtype == torch.float64) else None,
                              sharded=shard is not None and shard[1] > 1,
                              resident=self._transport in ("resident", "pull", "push", "none"),
                              force_tma=self._transport == "push")
        self._res_segs = {}
        self.flops = self.gemm.flops
        self.gemm_bytes = self.gemm.bytes

    def run(self, a, b):
        from .UniTensor import UniTensor
        dt = a.dtype
        if b.dtype != dt:
            raise RuntimeError("Contract(a,b): dtype mismatch %s vs %s" % (dt, b.dtype))
        with _lib.on_device(a.device):
            res_ok = self.shard is not None and self.shard[1] > 1 and self._transport in ("resident", "pull", "push")
            A = a._arena_tensor(resident_ok=res_ok)
            B = b._arena_tensor(resident_ok=res_ok)
            sharded = self.shard is not None and self.shard[1] > 1
            if sharded and self._transport in ("resident", "pull", "push"):
                return self._run_resident(a, b)
            fused = sharded and self._transport == "fused" and self.gemm.fusable()
            need_zero = (not self.all_matched) or (self.shard is not None and self._transport == "none")
            # every buffer is allocated before the first launch, so that the GEMM launch follows the ~10 us
            # relayout kernel without host work in between
            Ap = torch.empty(self.rel_a.packed.size, device=A.device, dtype=dt) if self.rel_a is not None else None
            Bp = torch.empty(self.rel_b.packed.size, device=B.device, dtype=dt) if self.rel_b is not None else None
            Cc = self.Lo.new_arena(dt, zero=need_zero and not fused)   # fused: overwritten by the copy-out
            if self.rel_a is not None:
                A = self.rel_a.run(A, Ap)
            if self.rel_b is not None:
                B = self.rel_b.run(B, Bp)
            if sharded:
                from . import parallel
                if self._transport in ("fused", "p2p") and self._stage is None:
                    self._stage = parallel.symm_staging(self, Cc.numel(), dt, Cc.device)
                if fused:
                    # ONE kernel: every output tile is stored into every rank's symmetric staging arena through
                    # NVLink peer memory while the remaining tiles are computed; barrier; local copy out
                    parallel.gemm_allgather(self._stage, self.gemm, A, B, Cc)
                elif self._transport in ("fused", "p2p"):
                    # own sectors -> symmetric staging arena; barrier; pull every slice into the result
                    stage = self._stage
                    self.gemm.run(A, B, stage.bufs[stage.turn])
                    parallel.gather_p2p(stage, Cc, self.ranges, self.shard[0])
                else:
                    self.gemm.run(A, B, Cc)
                    if self._transport == "nccl":
                        parallel.gather_ranges(Cc, self.ranges, self.shard[0])
            else:
                # the GEMM's set-up overlaps the tail of the relayout that precedes it (programmatic dependent launch)
                self.gemm.run(A, B, Cc, pdl=(self.rel_a is not None or self.rel_b is not None) and _PDL)
        return self._make_out(Cc)


def _resident_segments(plan, which, rel, t):
    """Segment table of relayout `rel` for a RESIDENT source tensor: the `source` column picks, per piece, the arena
    it is read from -- 0 = this rank's, k > 0 = the NVLink peer mapping of the k-th foreign owner.  Cached per
    (owner, valid) state of the operand.  Returns (table, foreign owners) or None when the plan has no segments."""
    if rel.d_seg is None:
        return None
    res = t._res
    key = (which, res.owner.tobytes(), res.valid.tobytes())
    hit = plan._res_segs.get(key)
    if hit is None:
        M = rel.M
        seg = rel.d_seg.clone()
        src = seg[:, 1].to(torch.int64)
        off = torch.from_numpy(np.asarray(M.arena_off, dtype=np.int64)).to(seg.device)
        sec = torch.clamp(torch.searchsorted(off, src, right=True) - 1, min=0)
        owner = torch.from_numpy(np.where(res.valid, -1, res.owner)).to(seg.device)[sec]      # -1: read locally
        owner = torch.where(src >= 0, owner, torch.full_like(owner, -1))
        foreign = sorted(int(x) for x in torch.unique(owner).tolist() if x >= 0)
        col = torch.zeros_like(owner)
        for k, o in enumerate(foreign):
            col[owner == o] = k + 1
        seg[:, 3] = col.to(torch.int32)
        hit = (seg.contiguous(), foreign)
        if len(plan._res_segs) > 8:
            plan._res_segs.pop(next(iter(plan._res_segs)))
        plan._res_segs[key] = hit
    return hit


def _run_resident(self, a, b):
    """Sector-resident Contract (parallel transport "resident" / "pull"): this rank relayouts and multiplies only
    the sectors it owns and keeps them; nothing is gathered ("pull": every rank then pulls the others' sectors
    into its own arena over NVLink -- a gathered result without a staging copy).  See parallel.py."""
    import weakref
    from . import parallel
    from .UniTensor import UniTensor
    dt = a.dtype
    rank, world = self.shard
    rw = parallel.resident_world(a.device)
    seq = rw.next_seq()
    A = a._arena_tensor(resident_ok=True)
    B = b._arena_tensor(resident_ok=True)
    need = [0] * world
    rel_args = {}
    for which, t, rel, secs in (("a", a, self.rel_a, self.sec_a), ("b", b, self.rel_b, self.sec_b)):
        res = t._res
        if res is None:
            continue
        if res.slot is None and not res.valid.all():
            # a resident operand outside peer-readable memory (the factors of a sharded Svd_truncate): when ANY rank
            # needs a sector it does not hold -- decided from the owner tables, identically on every rank -- the
            # tensor is completed by a collective; otherwise everybody reads what it has
            k = 1 if which == "a" else 2
            if rel is not None or any(int(res.owner[m[k]]) not in (-1, int(self.sector_owner[m[0]])) for m in self.matched_all):
                parallel.gather_collective(t)
        if res.slot is not None:
            res.slot.last_read = max(res.slot.last_read, seq)
        if rel is None:
            parallel.pull_sectors(t, secs)              # read in place: foreign sectors come once, then stay
            continue
        if rel.d_seg is None and rel.use_index and rel.d_index is None:
            rel._build_index()                          # (first use of the plan)
        hit = _resident_segments(self, which, rel, t)
        if hit is None or len(hit[1]) > 7:
            parallel.pull_sectors(t, range(len(res.owner)))    # no segment form: bring the whole tensor here
            continue
        for o in hit[1]:
            need[o] = max(need[o], res.seq)
        rel_args[which] = hit
    pool = self._pool
    if pool is False:          # first run: the ring of result arenas belongs to the plan (freed with it)
        pool = self._pool = rw.new_pool(self.Lo.arena_elems + ARENA_ALIGN, dt, a.device)
    slot, war = pool.acquire() if pool is not None else (None, 0)
    need = [max(n, war) if r != rank else 0 for r, n in enumerate(need)]
    rw.wait(need)
    Ap = torch.empty(self.rel_a.packed.size, device=A.device, dtype=dt) if self.rel_a is not None else None
    Bp = torch.empty(self.rel_b.packed.size, device=B.device, dtype=dt) if self.rel_b is not None else None
    for which, t, rel, src, dst in (("a", a, self.rel_a, A, Ap), ("b", b, self.rel_b, B, Bp)):
        if rel is None:
            continue
        hit = rel_args.get(which)
        if hit is None:
            rel.run(src, dst)
        else:
            ptrs = [src.data_ptr()] + [t._res.slot.peer_ptr(o) for o in hit[1]]
            rel.run_segments(ptrs, dst, _lib.dtype_code(dt), seg=hit[0], nseg=int(hit[0].shape[0]))
    Cc = slot.tensor() if slot is not None else self.Lo.new_arena(dt, zero=not self.all_matched)
    # (decided identically on every rank: fp64 symmetric sectors are always aligned, so a rank with work runs the TMA
    # kernel; a rank without any only signals)
    pushing = self._transport == "push" and slot is not None and dt == torch.float64 \
        and (getattr(self.gemm, "tma", False) or self.gemm.nprob == 0)
    sig = (rw.sig_ptrs, seq, rw.done_ctr)
    if pushing:
        # fused GEMM -> all-gather: finished tiles are stored into every other rank's ring slot as well
        import ctypes as C
        ptrs = slot.pool.push_args.get(slot.index)
        if ptrs is None:
            ptrs = slot.pool.push_args[slot.index] = (C.c_void_p * (world - 1))(
                *[slot.peer_ptr(r) for r in range(world) if r != rank])
        sig = sig + (ptrs,)
    self.gemm.run(Ap if Ap is not None else A, Bp if Bp is not None else B, Cc,
                  pdl=(self.rel_a is not None or self.rel_b is not None) and _PDL, signal=sig)
    out = self._make_out(Cc)
    out._res = parallel.Residency(world, rank, self.sector_owner, seq, slot, self._own_valid)
    if slot is not None:
        slot.holder = weakref.ref(out)
    if self._transport == "pull" or (self._transport == "push" and not pushing):
        if slot is not None:
            parallel.pull_sectors(out, range(self.Lo.nsec))
        else:
            parallel.gather_collective(out)
    elif pushing:
        # the whole result is here once every peer has published this operation (its pushes precede its signal)
        rw.wait([seq if r != rank else 0 for r in range(world)])
        out._res.valid[:] = True
        slot.last_read = max(slot.last_read, seq)
    return out


def _make_out(self, Cc):
    """The result tensor around arena Cc (metadata replayed from a template built on the plan's first run)."""
    from .UniTensor import UniTensor
    tpl = self._out_tpl
    if tpl is None:
        tpl = UniTensor._from_plan(self.out_bonds, self.out_labels, self.out_braket, self.mem_rowrank, self.Lo, None)
        if self.out_rowrank != self.mem_rowrank:
            tpl.rowrank = self.out_rowrank
            tpl._bq = None
            tpl._check_braket()
        self._out_tpl = tpl
    return UniTensor._from_template(tpl, Cc)


def _contract_plan_bytes(self):
    """The GEMM group's tables and workspace plus the peer-source segment tables (the relayouts are entries of their
    own cache; result rings are result storage, not plan memory)."""
    n = self.gemm.plan_bytes() if getattr(self, "gemm", None) is not None else 0
    for seg, _ in getattr(self, "_res_segs", {}).values():
        n += int(seg.numel()) * seg.element_size()
    return n


ContractPlan.plan_bytes = _contract_plan_bytes
ContractPlan._make_out = _make_out
ContractPlan._out_tpl = None
ContractPlan._pool = False


ContractPlan._run_resident = _run_resident


def contract_key(a, b):
    # layouts are cached objects: their identity stands for their (large, nested) signature; a plan keeps the
    # operand layouts it was built for alive (ContractPlan._keep), so an id cannot be recycled under a live key
    return (id(a._layout), a._mapper.tobytes(), a.labels.tobytes(), a.rowrank,
            id(b._layout), b._mapper.tobytes(), b.labels.tobytes(), b.rowrank,
            a.dtype, a.braket.tobytes(), b.braket.tobytes())


def get_contract_plan(a, b):
    from . import parallel
    shard = parallel.current_shard()
    if shard is not None and _SHARD_MIN_FLOP > 0:
        # small contractions are not worth a launch per rank plus bookkeeping: below the threshold every rank
        # computes the whole (replicated) result; the estimate is the dense-equivalent row x col x k of the operands
        if 2.0 * a._layout.nnz * max(b._layout.nnz, 1) ** 0.5 < _SHARD_MIN_FLOP:
            shard = None
    rk = None
    if shard and a._res is not None and parallel._STATE["transport"] in ("resident", "pull", "push"):
        rk = a._res.key()          # ownership may propagate from a resident first operand
    key = (contract_key(a, b), shard, parallel._STATE["transport"] if shard else None, rk)
    pl = _CONTRACTS.get(key)
    if pl is None:
        pl = ContractPlan(a, b, shard)
        _CONTRACTS[key] = pl
    return pl
