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

_STATE = {"group": None, "enabled": False}


def enable(group=None):
    if not dist.is_initialized():
        raise RuntimeError("tor10_b200.parallel.enable(): torch.distributed is not initialised")
    _STATE["group"] = group
    _STATE["enabled"] = True


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


def gather_ranges(arena, ranges, rank, group=None):
    """In-place variable-size all-gather: after the call every rank holds all slices.  NCCL: one
    coalesced group of broadcasts (torch's uneven all_gather); gloo (CPU tests): one broadcast per owner."""
    group = _STATE["group"] if group is None else group
    slices = [arena[lo:hi] for lo, hi in ranges]
    if dist.get_backend(group) == "nccl":
        dist.all_gather(slices, slices[rank], group=group)
    else:
        for r, sl in enumerate(slices):
            if sl.numel():
                dist.broadcast(sl, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return arena
