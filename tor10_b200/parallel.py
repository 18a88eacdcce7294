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

_STATE = {"group": None, "enabled": False, "transport": "auto"}


def enable(group=None, transport="auto"):
    """transport: "fused" = the grouped GEMM stores every output tile into all ranks' symmetric staging arenas
                           (NVLink peer stores in its epilogue, overlapped with the math) + one device barrier
                           + a local copy out; plans whose kernel cannot fuse (fp32) take the p2p path,
                  "p2p"   = symmetric-memory staging + device barrier + one pull kernel over NVLink peer memory,
                  "nccl"  = all_gather_into_tensor through a staging buffer,
                  "none"  = results stay sharded: every rank holds only its own sectors (others are zero),
                  "auto"  = fused on NCCL process groups when symmetric memory is available, else nccl."""
    if not dist.is_initialized():
        raise RuntimeError("tor10_b200.parallel.enable(): torch.distributed is not initialised")
    if transport not in ("auto", "fused", "p2p", "nccl", "none"):
        raise ValueError("transport must be auto | fused | p2p | nccl | none")
    _STATE["group"] = group
    _STATE["enabled"] = True
    _STATE["transport"] = transport


def disable():
    _STATE["enabled"] = False


def current_shard():
    """(rank, world) when sector sharding is on, else None."""
    if not _STATE["enabled"]:
        return None
    g = _STATE["group"]
    world = dist.get_world_size(g)
    if world == 1:
        return None
    return (dist.get_rank(g), world)


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
        return "fused"         # GEMM with the all-gather in its epilogue; plans that cannot fuse pull (p2p)
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
