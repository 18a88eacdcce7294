"""Multi-GPU: independent symmetry sectors partitioned over the ranks of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink).  The reference has no counterpart (no
multi-device code at all).  A symmetric Contract has no cross-sector reduction: output sector q needs only
A_q and B_q (UniTensor.py:3727-3741), so every rank relayouts and multiplies ONLY the sectors it owns and
the single collective is the gather of finished sectors.  Ownership is a contiguous run of output sectors
(balanced by flops), which makes each rank's result one contiguous slice of the output arena: the gather
is one variable-size all-gather written in place, no packing kernels.

    tor10_b200.parallel.enable()          # after dist.init_process_group; Contract() is now sharded
    tor10_b200.parallel.disable()
"""
import numpy as np
import torch
import torch.distributed as dist

_STATE = {"group": None, "enabled": False, "transport": "auto", "shard": None}


def enable(group=None, transport="auto"):
    """transport: "fused" = the grouped GEMM stores every output tile into all ranks' symmetric staging arenas
                           (NVLink peer stores in its epilogue, overlapped with the math) + one device barrier
                           + a local copy out; plans whose kernel cannot fuse (fp32) take the p2p path,
                  "p2p"   = symmetric-memory staging + device barrier + one pull kernel over NVLink peer memory,
                  "nccl"  = all_gather_into_tensor through a staging buffer,
                  "none"  = results stay sharded: every rank holds only its own sectors (others are zero),
                  "resident" = results stay sector-RESIDENT: every rank keeps the sectors it computed, the tensor
                           carries an owner table, later contractions read it in place / fetch foreign sectors
                           from their owners over NVLink, and anything that needs the whole tensor gathers it
                           on demand (see the second half of this module),
                  "pull"  = resident, then every rank pulls the other ranks' sectors at once (gathered result,
                           no staging copy),
                  "push"  = fused GEMM -> all-gather on the TMA kernel: finished tiles are stored from the
                           accumulator registers into every rank's ring slot over NVLink while the remaining pieces
                           compute; a rank holds the whole result once it has acquired every peer's progress word
                           (gathered result, no staging, no copy),
                  "auto"  = fused on NCCL process groups when symmetric memory is available, else nccl."""
    if not dist.is_initialized():
        raise RuntimeError("tor10_b200.parallel.enable(): torch.distributed is not initialised")
    if transport not in ("auto", "fused", "p2p", "nccl", "none", "resident", "pull", "push"):
        raise ValueError("transport must be auto | fused | p2p | nccl | none | resident | pull | push")
    _STATE["group"] = group
    _STATE["enabled"] = True
    _STATE["transport"] = transport
    world = dist.get_world_size(group)
    _STATE["shard"] = None if world == 1 else (dist.get_rank(group), world)


def disable():
    _STATE["enabled"] = False


def current_shard():
    """(rank, world) when sector sharding is on, else None."""
    if not _STATE["enabled"]:
        return None
    return _STATE["shard"]


def partition_contiguous(cost, parts):
    """Split the sequence `cost` into `parts` contiguous runs minimising the largest run sum (linear
    partition, exact by binary search on the bottleneck).  Returns bounds[0..parts] (run r = [b[r], b[r+1]))."""
    cost = np.asarray(cost, dtype=np.float64)
    n = len(cost)

    def runs_needed(cap):
        cnt, acc = 1, 0.0
        for c in cost:
            if acc + c > cap and acc > 0:
                cnt, acc = cnt + 1, 0.0
            acc += c
        return cnt

    lo, hi = (cost.max() if n else 0.0), cost.sum()
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if runs_needed(mid) <= parts:
            hi = mid
        else:
            lo = mid
    bounds, acc = [0], 0.0
    for i, c in enumerate(cost):
        if acc + c > hi * (1 + 1e-12) and acc > 0 and len(bounds) < parts:
            bounds.append(i)
            acc = 0.0
        acc += c
    while len(bounds) < parts:
        bounds.append(n)
    bounds.append(n)
    return [int(b) for b in bounds]


def partition_lpt(cost, parts):
    """Owner of every item by longest-processing-time-first: heaviest item to the least loaded part.  Used
    when ownership need not be contiguous (fused GEMM -> all-gather stores tiles, not arena slices)."""
    cost = np.asarray(cost, dtype=np.float64)
    owner = np.zeros(len(cost), dtype=np.int64)
    load = np.zeros(parts)
    for i in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += cost[i]
    return owner


def arena_ranges(layout, bounds):
    """Arena element range [lo, hi) that holds sectors bounds[r]..bounds[r+1]-1, for every rank r."""
    out = []
    for r in range(len(bounds) - 1):
        s0, s1 = bounds[r], bounds[r + 1]
        if s0 >= s1:
            out.append((0, 0))
            continue
        lo = int(layout.arena_off[s0])
        hi = int(layout.arena_off[s1]) if s1 < layout.nsec else int(layout.arena_elems)
        out.append((lo, hi))
    return out


_STAGE = {}
_SYMM = {}
import os as _os


class _SymmStaging:
    """Two symmetric-memory staging arenas per plan, used alternately (a rank may run one step ahead of its
    peers, never two), plus -- for the fused GEMM -> all-gather -- one symmetric uint64 flag array per arena
    (slot r = "rank r has stored all its tiles of step s") and a local arrival counter."""

    def __init__(self, n, dtype, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bufs, self.hdls, self.flags, self.flag_hdls = [], [], [], []
        for _ in range(2):
            t = symm_mem.empty(n, dtype=dtype, device=device)
            t.zero_()     # sectors no rank ever writes (unmatched, padding) must read as zero
            f = symm_mem.empty(self.world, dtype=torch.int64, device=device)
            f.zero_()
            torch.cuda.synchronize(device)      # zeroed before any peer can reach them
            self.hdls.append(symm_mem.rendezvous(t, group=grp))
            self.flag_hdls.append(symm_mem.rendezvous(f, group=grp))
            self.bufs.append(t)
            self.flags.append(f)
        self.turn = 0
        self.args = {}
        self.n = n
        self.steps = [0, 0]
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.peer_ptrs = [[int(h.get_buffer(r, (n,), dtype).data_ptr()) for r in range(self.world)]
                          for h in self.hdls]
        # address of THIS rank's slot inside rank r's flag array
        self.peer_flag_ptrs = [[int(h.get_buffer(r, (self.world,), torch.int64).data_ptr()) + 8 * self.rank
                                for r in range(self.world)] for h in self.flag_hdls]
        dist.barrier(group=group)


def symm_staging(owner, n, dtype, device, group=None):
    """Staging arenas are private to one plan (`owner`): regions a plan never writes stay zero."""
    group = _STATE["group"] if group is None else group
    key = (getattr(owner, "serial", id(owner)), device.index, dtype, n)
    st = _SYMM.get(key)
    if st is None:
        st = _SymmStaging(n, dtype, device, group)
        _SYMM[key] = st
    return st


def transport_for(group=None):
    group = _STATE["group"] if group is None else group
    t = _STATE["transport"]
    if t != "auto":
        return t
    if dist.get_backend(group) != "nccl":
        return "nccl"          # (gloo tests: broadcast path inside gather_ranges)
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        # results stay sector-resident (gathered on demand); TOR10_TRANSPORT_AUTO=fused restores the round-1 default
        # (GEMM with the all-gather in its epilogue; plans that cannot fuse pull)
        return _os.environ.get("TOR10_TRANSPORT_AUTO", "resident")
    except Exception:
        return "nccl"


def gemm_allgather(stage, gemm, A, B, arena):
    """Fused path, two launches: (1) the grouped GEMM stores every output tile into all ranks' staging arenas
    (NVLink peer stores overlapped with the math) and its last CTA raises this rank's flag on every peer
    (t10_ggemm_allgather); (2) t10_wait_copy acquires all ranks' flags and copies the completed staging arena
    into the result arena (a local HBM copy).  No host-side or NCCL synchronisation on the data path."""
    import ctypes as C
    from . import _lib
    k = stage.turn
    args = stage.args.get(("push", k))
    if args is None:
        args = ((C.c_void_p * stage.world)(*stage.peer_ptrs[k]), (C.c_void_p * stage.world)(*stage.peer_flag_ptrs[k]))
        stage.args[("push", k)] = args
    stage.steps[k] += 1
    step = stage.steps[k]
    gemm.run_allgather(A, B, args[0], args[1], step, stage.done_ctr, self_index=stage.rank)
    n = arena.numel() // 4 * 4      # whole 16-byte vectors (the arena ends in >= 16 elements of padding)
    _lib.check(_lib.load().t10_wait_copy(_lib.dtype_code(arena.dtype), arena.data_ptr(), stage.bufs[k].data_ptr(),
                                         n, stage.flags[k].data_ptr(), stage.world, step,
                                         stage.err.data_ptr(), _lib.stream_ptr()))
    stage.turn ^= 1
    return arena


def fused_status(stage):
    """True if a t10_wait_copy of this staging set ever timed out waiting for a peer (synchronises)."""
    return bool(stage.err.item())


def gather_p2p(stage, arena, ranges, rank):
    """`stage` = this step's symmetric staging arena (GEMM output of every rank's own sectors).  Device
    barrier, then one kernel pulls the peers' slices over NVLink and copies the local slice."""
    import ctypes as C
    from . import _lib
    k = stage.turn
    stage.hdls[k].barrier(channel=0)
    args = stage.args.get(k)
    if args is None:   # ctypes argument arrays are built once per staging buffer
        # start with the next rank's slice and wrap around: the ranks then read from different peers at any
        # moment instead of all pulling rank 0's slice first
        order = [(rank + 1 + i) % len(ranges) for i in range(len(ranges))]
        segs = [(stage.peer_ptrs[k][r], ranges[r][0], ranges[r][1]) for r in order if ranges[r][1] > ranges[r][0]]
        n = len(segs)
        args = (n, (C.c_void_p * n)(*[p for p, _, _ in segs]), (C.c_int64 * n)(*[a for _, a, _ in segs]),
                (C.c_int64 * n)(*[b for _, _, b in segs]), _lib.dtype_code(arena.dtype))
        stage.args[k] = args
    n, src, lo, hi, dcode = args
    _lib.check(_lib.load().t10_gather_peers(dcode, arena.data_ptr(), n, src, lo, hi, _lib.stream_ptr()))
    stage.turn ^= 1
    return arena


def gather_ranges(arena, ranges, rank, group=None):
    """In-place variable-size all-gather: after the call every rank holds all slices.
    NCCL: ONE equal-size `all_gather_into_tensor` through a persistent staging buffer (every rank sends
    max_len elements starting at its slice; the tail beyond a shorter slice is ignored), then the peers'
    slices are copied to their place -- measured 2.5x cheaper per call than torch's uneven `all_gather`
    (a coalesced group of broadcasts) at 25 MB.  gloo (CPU tests): one broadcast per owner."""
    group = _STATE["group"] if group is None else group
    world = len(ranges)
    if dist.get_backend(group) != "nccl":
        for r, (lo, hi) in enumerate(ranges):
            if hi > lo:
                dist.broadcast(arena[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r,
                               group=group)
        return arena
    max_len = max(hi - lo for lo, hi in ranges)
    if max_len == 0:
        return arena
    lo, hi = ranges[rank]
    key = (arena.device.index, arena.dtype, world, max_len)
    stage = _STAGE.get(key)
    if stage is None:
        stage = torch.empty(world * max_len, device=arena.device, dtype=arena.dtype)
        _STAGE[key] = stage
    if lo + max_len <= arena.numel():
        send = arena[lo:lo + max_len]
    else:  # last slice is shorter than max_len and sits at the end of the arena
        send = stage.new_empty(max_len)
        send[:hi - lo].copy_(arena[lo:hi])
    dist.all_gather_into_tensor(stage, send, group=group)
    for r, (plo, phi) in enumerate(ranges):
        if r != rank and phi > plo:
            arena[plo:phi].copy_(stage[r * max_len:r * max_len + (phi - plo)])
    return arena


# ======================================================================================================
# Sector-RESIDENT tensors ("preferred: keep tensors sector-sharded across consecutive
# contractions; ownership propagates").
#
# transport "resident": Contract() returns a tensor of which every rank holds only the sectors it computed.
# Nothing is gathered.  The tensor carries an owner table; a later Contract adopts it (output sector q lives
# where A's row sector q lives), reads what it owns in place and fetches the few foreign sectors it needs
# straight out of the owner's arena through NVLink peer mappings (inside the relayout copy, or by a pull
# kernel) -- no collective on the data path.  Anything that needs the whole tensor (`.Storage`, GetBlock,
# arithmetic, pickling, printing) gathers it on demand, once.
#
# Synchronisation is a single monotone PROGRESS WORD per rank, in symmetric memory (peers can read it):
#   * every resident operation has a sequence number (all ranks issue the same operations in the same order:
#     SPMD, the contract of torch.distributed itself); the kernel that completes operation s on rank r stores s
#     into rank r's OWN word (a local st.release.sys, t10_ggemm_tma / t10_signal_flags) -- measured: storing it
#     into the peers' copies instead cost 12.5 us per Contract at 2 GPUs, the latency of one NVLink store;
#   * read-after-write: before a kernel reads sectors rank p produced in operation s it acquires word p >= s,
#     polling it REMOTELY through the NVLink mapping -- only a reader that needs foreign data pays for the link;
#   * write-after-read: result arenas come from a small ring per plan (symmetric memory, so peers can read
#     them); before a slot is overwritten the ranks acquire every peer's word >= the last operation that read it.
# Waits are one-warp kernels (t10_wait_flags); they are skipped when the host already knows the value was
# acquired on the stream.  A bounded wait that expires traps (never continue on
# a timeout).
# ======================================================================================================
import os as _os
import weakref as _weakref

_RES = {"world": None}
WAIT_TIMEOUT_MS = int(_os.environ.get("TOR10_WAIT_TIMEOUT_MS", "20000"))
POOL_DEPTH = int(_os.environ.get("TOR10_RESIDENT_POOL", "4"))


class Residency:
    """Where the sectors of a resident tensor live.  owner[s] = rank that computed sector s (-1: nobody, the
    sector is zero everywhere); valid[s] = this rank's arena holds sector s; seq = the operation that produced
    it (read-after-write bound on the owners' progress words)."""
    __slots__ = ("world", "rank", "owner", "valid", "seq", "slot", "__weakref__")

    def __init__(self, world, rank, owner, seq, slot, own_valid=None):
        self.world, self.rank = world, rank
        self.owner = owner                     # (shared with the plan: never modified)
        self.valid = ((np.asarray(owner) == rank) | (np.asarray(owner) < 0)) if own_valid is None else own_valid.copy()
        self.seq, self.slot = seq, slot

    def key(self):
        return (self.world, self.owner.tobytes())


class _Slot:
    __slots__ = ("pool", "index", "holder", "last_read", "pull_args")

    def __init__(self, pool, index):
        self.pool, self.index, self.holder, self.last_read = pool, index, None, 0
        self.pull_args = {}          # cached argument arrays of pull_sectors (they hold this slot's peer addresses)

    def tensor(self):
        return self.pool.bufs[self.index]

    def peer_ptr(self, r):
        return self.pool.peer_ptrs[self.index][r]


class ArenaPool:
    """Ring of POOL_DEPTH result arenas of one plan in ONE symmetric-memory allocation (peers read resident
    sectors out of them).  Regions no rank ever writes (unmatched sectors, padding) stay zero."""

    def __init__(self, rw, n, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.rw, self.n, self.dtype = rw, int(n), dtype
        self.stride = (self.n + 63) // 64 * 64
        esz = torch.empty(0, dtype=dtype).element_size()
        depth = POOL_DEPTH if self.stride * esz <= (256 << 20) else 2     # large results: two slots
        self.depth = depth
        whole = symm_mem.empty(depth * self.stride, dtype=dtype, device=device)
        whole.zero_()
        torch.cuda.synchronize(device)
        hdl = symm_mem.rendezvous(whole, group=rw.group_or_world())
        self.whole, self.hdl = whole, hdl
        self.bufs = [whole[k * self.stride:k * self.stride + self.n] for k in range(depth)]
        base = [int(hdl.get_buffer(r, (depth * self.stride,), dtype).data_ptr()) for r in range(rw.world)]
        self.peer_ptrs = [[base[r] + k * self.stride * esz for r in range(rw.world)] for k in range(depth)]
        self.slots = [_Slot(self, k) for k in range(depth)]
        self.push_args = {}
        self.turn = 0
        dist.barrier(group=rw.group)

    def acquire(self):
        """Next slot of the ring.  A tensor still living in it is gathered out first (every rank does so at the
        same program point).  Returns (slot, write-after-read bound)."""
        slot = self.slots[self.turn]
        self.turn = (self.turn + 1) % self.depth
        old = slot.holder() if slot.holder is not None else None
        if old is not None and getattr(old, "_res", None) is not None and old._res.slot is slot:
            old._materialize()
        slot.holder = None
        return slot, slot.last_read


class ResidentWorld:
    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.progress = symm_mem.empty(8, dtype=torch.int64, device=device)      # word 0 = this rank's progress
        self.progress.zero_()
        torch.cuda.synchronize(device)
        hdl = symm_mem.rendezvous(self.progress, group=self.group_or_world())
        self.hdl = hdl
        import ctypes as C
        # the producer stores into its OWN word only; readers poll the word of the rank they depend on (a peer's
        # through its NVLink mapping), so a rank that nobody waits for generates no link traffic at all
        self.word_ptrs = (C.c_void_p * self.world)(*[int(hdl.get_buffer(r, (8,), torch.int64).data_ptr())
                                                     for r in range(self.world)])
        self.sig_ptrs = (C.c_void_p * 1)(int(self.progress.data_ptr()))
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.seq = 0
        self.acquired = {}          # stream -> per-rank value known to have been acquired on it
        dist.barrier(group=group)

    def group_or_world(self):
        return self.group if self.group is not None else dist.group.WORLD

    def next_seq(self):
        self.seq += 1
        return self.seq

    def wait(self, need):
        """Acquire progress[r] >= need[r] on the current stream (a one-warp kernel; skipped when known)."""
        if not any(need):
            return False
        import ctypes as C
        from . import _lib
        st = _lib.stream_ptr()
        have = self.acquired.setdefault(st, [0] * self.world)
        have[self.rank] = max(have[self.rank], self.seq)      # this rank's own operations are stream-ordered
        need = [int(x) for x in need]
        if all(n <= h for n, h in zip(need, have)):
            return False
        arr = (C.c_uint64 * self.world)(*[max(n, 0) for n in need])
        _lib.check(_lib.load().t10_wait_flags(self.world, self.word_ptrs, arr, WAIT_TIMEOUT_MS,
                                              self.err.data_ptr(), st))
        for r in range(self.world):
            have[r] = max(have[r], need[r])
        return True

    def signal(self, value):
        """Publish `value` as this rank's progress from a one-warp kernel (after read-only operations)."""
        from . import _lib
        _lib.check(_lib.load().t10_signal_flags(1, self.sig_ptrs, int(value), _lib.stream_ptr()))

    def new_pool(self, n, dtype, device):
        """A ring of symmetric result arenas for one plan (owned by the plan: evicting the plan frees it), or None
        when symmetric memory of that size cannot be had (the plan's results then live in private memory: still
        sector-resident, completed by a collective on demand)."""
        try:
            p = ArenaPool(self, n, dtype, device)
        except RuntimeError:
            p = None
        ok = torch.tensor([1 if p is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)            # every rank takes the same decision
        if int(ok.item()) == 0:
            p = None
        return p


def resident_world(device=None):
    rw = _RES["world"]
    if rw is None:
        if dist.get_backend(_STATE["group"]) != "nccl":
            raise RuntimeError("tor10_b200.parallel: resident tensors need NVLink peer memory (NCCL process group)")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if dist.get_world_size(_STATE["group"]) > 8:
            raise RuntimeError("tor10_b200.parallel: resident tensors span the (at most 8) GPUs of ONE box; use a "
                               "process group of that size")
        rw = ResidentWorld(_STATE["group"], device)
        _RES["world"] = rw
    return rw


def reset_resident():
    _RES["world"] = None


def inherit_owner(a_owner, matched, nsec_out, cost, world):
    """Ownership of the output sectors of a contraction whose first operand is resident and read in place: output
    sector q lives where A's row sector q lives; sectors A's table does not place (owner -1: A's
    sector is zero) and unmatched ones go to nobody.  Pure function (tested on the CPU)."""
    owner = np.full(nsec_out, -1, dtype=np.int64)
    for ob, ab, _ in matched:
        owner[ob] = a_owner[ab]
    return owner


def fetch_plan(layout, res, needed):
    """Which of the `needed` sectors this rank must pull, grouped by owner: {owner: [(lo, hi), ...]} arena ranges.
    Pure function (tested on the CPU)."""
    out = {}
    for s in needed:
        if res.valid[s]:
            continue
        o = int(res.owner[s])
        lo = int(layout.arena_off[s])
        hi = lo + int(layout.rows[s]) * int(layout.ld[s])
        hi = (hi + 3) // 4 * 4          # whole 16-byte vectors in fp32 too (sector bases are 16-element aligned)
        out.setdefault(o, []).append((lo, hi))
    return out


def pull_sectors(t, needed):
    """Make the `needed` sectors of resident tensor t valid in this rank's arena: one pull kernel over the owners'
    NVLink peer mappings (after acquiring their progress).  No-op for sectors already here."""
    import ctypes as C
    from . import _lib
    res = t._res
    plan = fetch_plan(t._layout, res, needed)
    if not plan:
        return
    rw = resident_world()
    if res.slot is None:
        raise RuntimeError("tor10_b200.parallel: this resident tensor no longer lives in peer-readable memory")
    need = [0] * rw.world
    for o in plan:
        need[o] = res.seq
    rw.wait(need)
    # one launch for (up to 64) slices: neighbouring sectors of one owner travel as one slice; the argument arrays are
    # cached ON the slot per request (a "pull" transport asks for the same sectors every step; the cache dies with
    # the slot whose peer addresses it holds)
    key = (tuple(needed) if not isinstance(needed, range) else ("range", needed.start, needed.stop, needed.step),
           res.owner.tobytes())
    cache = res.slot.pull_args
    args = cache.get(key)
    if args is None:
        segs = []
        for o, rs in sorted(plan.items(), key=lambda kv: (kv[0] - res.rank) % res.world):   # start with the next rank
            for lo, hi in sorted(rs):
                if segs and segs[-1][0] == o and segs[-1][2] >= lo:
                    segs[-1][2] = max(segs[-1][2], hi)
                else:
                    segs.append([o, lo, hi])
        args = []
        for i in range(0, len(segs), 64):
            part = segs[i:i + 64]
            n = len(part)
            args.append((n, (C.c_void_p * n)(*[res.slot.peer_ptr(o) for o, _, _ in part]),
                         (C.c_int64 * n)(*[a for _, a, _ in part]), (C.c_int64 * n)(*[b for _, _, b in part])))
        if len(cache) > 64:
            cache.clear()
        cache[key] = args
    arena = t._arena
    for n, ptrs, los, his in args:
        _lib.check(_lib.load().t10_gather_peers(_lib.dtype_code(arena.dtype), arena.data_ptr(), n, ptrs, los, his,
                                                _lib.stream_ptr()))
    for s in needed:
        res.valid[s] = True
    # the peers read THIS rank's slot the same way: they are done once they have published their next operation
    res.slot.last_read = max(res.slot.last_read, rw.seq + 1)


def gather_collective(t):
    """Complete a resident tensor that does NOT live in peer-readable memory (results assembled by torch ops, e.g. the
    factors of a sharded Svd_truncate): one broadcast per owner and contiguous run of its sectors.  A collective:
    every rank must call it at the same program point."""
    res = t._res
    group = _STATE["group"]
    arena = t._arena
    L = t._layout
    run_owner, run_lo, run_hi = None, 0, 0
    runs = []
    for s in range(len(res.owner)):
        o = int(res.owner[s])
        lo = int(L.arena_off[s])
        hi = lo + int(L.rows[s]) * int(L.ld[s])
        if o == run_owner and o >= 0:
            run_hi = hi
        else:
            if run_owner is not None and run_owner >= 0:
                runs.append((run_owner, run_lo, run_hi))
            run_owner, run_lo, run_hi = o, lo, hi
    if run_owner is not None and run_owner >= 0:
        runs.append((run_owner, run_lo, run_hi))
    for o, lo, hi in runs:
        dist.broadcast(arena[lo:hi], src=dist.get_global_rank(group, o) if group is not None else o, group=group)
    res.valid[:] = True


def materialize(t):
    """Gather a resident tensor on demand: pull every foreign sector, publish the read (so the owners may reuse
    their slots), move the tensor into private memory.  Every rank that holds the tensor must call this at the
    same program point if it calls it at all (SPMD); ranks that never look at the whole tensor never pay for it."""
    res = t._res
    if res is None:
        return
    rw = resident_world()
    if res.slot is not None:
        pull_sectors(t, [s for s in range(len(res.owner)) if not res.valid[s]])
        seq = rw.next_seq()
        res.slot.last_read = max(res.slot.last_read, seq)
        t._arena = t._arena.clone()      # BEFORE the signal: once it is out, a peer may overwrite the slot ("push")
        rw.signal(seq)
        if res.slot.holder is not None and res.slot.holder() is t:
            res.slot.holder = None
    elif res.world > 1:
        gather_collective(t)
    t._res = None
    t._blocks = None
