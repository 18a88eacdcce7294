"""CUDA-graph replay of a fixed sequence of library calls (B200: launch-bound inner loops belong in a graph).

An algorithm step whose tensors keep their shapes -- an iTEBD / TEBD update at fixed bond dimension: ten small
Contracts, one small SVD, a handful of element-wise kernels, ~45 launches of 2-5 us each -- is bound by the host,
not by the GPU (reference: iTEBD.py:75-136 spends its time in the Python interpreter).  `StaticLoop` runs the step
function a few times eagerly (plans are built, caches warm, the storage forms of the state settle), captures ONE
call into a CUDA graph whose last nodes copy the new state back into the tensors it started from, and then replays
it: one graph launch per step.  Everything the library launches is capturable (no host synchronisation, no
allocation outside torch's graph pool); what is not (`.item()`, printing, a cuSOLVER SVD of a matrix too large for
the one-CTA kernel) belongs outside the step function.

    loop = StaticLoop(step, state=(A, B, la, lb), consts=(H, eH, chi))     # step(*state, *consts) -> (*new_state, *outputs)
    for i in range(n):
        outputs = loop()             # state tensors now hold step i's result; outputs are device tensors
"""
import torch

from . import _plan
from .UniTensor import UniTensor


def _storage_of(t):
    return t._dense if isinstance(t, UniTensor) else t


def _copy_state_(dst, src):
    """dst <- src, in place (dense UniTensors: storage and, when present, the diagonal hint)."""
    for d, s in zip(dst, src):
        if isinstance(d, UniTensor):
            if d.is_symm or s.is_symm:
                raise TypeError("StaticLoop: symmetric state tensors are not supported yet")
            if d.is_diag != s.is_diag or tuple(d._dense.shape) != tuple(s._dense.shape) or \
                    (d._diag_hint is None) != (s._diag_hint is None):
                raise RuntimeError("StaticLoop: the state changed its storage form between steps")
            hint = d._diag_hint
            d._dense.copy_(s._dense)             # (in place: the property setter is not involved)
            if hint is not None:
                hint.copy_(s._diag_hint)
            d.labels = s.labels.copy()
        else:
            d.copy_(s)


class StaticLoop:
    def __init__(self, step, state, consts=(), warmup=3):
        self.step, self.consts = step, tuple(consts)
        n = len(state)
        state = tuple(state)
        for _ in range(max(int(warmup), 1)):       # eager: builds every plan, lets the storage forms settle
            out = step(*state, *self.consts)
            state = tuple(out[:n])
        self.state = state
        self.n = n
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = step(*self.state, *self.consts)
            _copy_state_(self.state, out[:n])
            self.outputs = tuple(_storage_of(o).clone() for o in out[n:])
        torch.cuda.synchronize()
        # the graph addresses the device tables of the plans it used by raw pointer: keep them alive whatever the
        # LRU caches evict later
        self._pins = _plan.pinned_objects()
        self.steps = 0

    def __call__(self):
        self.graph.replay()
        self.steps += 1
        return self.outputs
