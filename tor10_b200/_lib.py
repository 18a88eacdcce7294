"""ctypes binding of the C-ABI in include/tor10_b200.h.

The product path has no CPU compute path: if the shared library is missing, or a compute
entry point is called without a CUDA device, this module raises -- it never substitutes torch
or numpy arithmetic.
"""
import ctypes as C
import os

import numpy as np
import torch  # noqa: F401  (loads libcudart / creates the CUDA context the kernels run in)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOR10_LIB") or os.path.join(HERE, "lib", "libtor10_b200.so")   # (TOR10_LIB: debug builds)

T10_F64, T10_F32 = 0, 1
T10_SYM_U1, T10_SYM_ZN = 0, 1
T10_MAX_RANK, T10_MAX_NSYM = 16, 8
ARENA_ALIGN = 16
RELAYOUT_TR, RELAYOUT_TC = 32, 128

STATUS = {1: "T10_ERR_INVALID", 2: "T10_ERR_CUDA", 3: "T10_ERR_KEYSPACE", 4: "T10_ERR_NOBLOCK",
          5: "T10_ERR_UNSUPPORTED"}


class T10Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


class NoValidBlock(TypeError):
    """UniTensor.py:415-417"""


class GemmProblem(C.Structure):
    _fields_ = [("a_off", C.c_int64), ("b_off", C.c_int64), ("c_off", C.c_int64),
                ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
                ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32)]


GEMM_PROBLEM_DTYPE = np.dtype([("a_off", "<i8"), ("b_off", "<i8"), ("c_off", "<i8"),
                               ("m", "<i4"), ("n", "<i4"), ("k", "<i4"),
                               ("lda", "<i4"), ("ldb", "<i4"), ("ldc", "<i4")])
assert GEMM_PROBLEM_DTYPE.itemsize == C.sizeof(GemmProblem) == 48
RELAYOUT_TILE_DTYPE = np.dtype([("dst", "<i8"), ("rbase", "<i4"), ("cbase", "<i4"), ("nr", "<i4"), ("nc", "<i4"),
                                ("dld", "<i4"), ("pad", "<i4")])
assert RELAYOUT_TILE_DTYPE.itemsize == 32

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64

# name -> argtypes ; every symbol include/tor10_b200.h declares
SIGNATURES = {
    "t10_version": [],
    "t10_last_error": [],
    "t10_device_props": [_i, _p, _p, _p, _p],
    "t10_fuse_qnums": [_i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "t10_blockmap_build": [_i, _i, _i, _p, _p, _p, _p, _p, _p, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "t10_unravel": [_l, _i, _p, _p, _p, _p],
    "t10_relayout_tables": [_l, _i, _p, _p, _p, _p, _p, _p, _p],
    "t10_relayout_blocks": [_i, _p, _p, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "t10_relayout_index": [_i, _p, _p, _l, _p, _p],
    "t10_gather_peers": [_i, _p, _i, _p, _p, _p, _p],
    "t10_permute_dense": [_i, _p, _p, _i, _p, _p, _p],
    "t10_diag_scale": [_i, _p, _p, _p, _l, _l, _l, _p],
    "t10_ggemm_tile_shape": [_i, _p, _p],
    "t10_ggemm_slots": [_i, _i, _p],
    "t10_ggemm_plan_tiles": [_i, _l, _p, _l, _l, _p, _p, _p],
    "t10_ggemm": [_i, _i, _p, _p, _p, _l, _p, _l, _p, _l, _p, _p, _p],
    "t10_ggemm_tf32_status": [_p],
    "t10_ggemm_set_trace": [_p],
    "t10_ggemm_streamk": [_p, _p, _p, _l, _p, _l, _p, _p, _l, _p, _p, _p, _i, _p],
    "t10_ggemm_streamk_allgather": [_p, _p, _i, _i, _p, _p, C.c_uint64, _p, _l, _p, _l, _p, _p, _l, _p, _p, _p, _i, _p],
    "t10_reduce_slices": [_i, _p, _p, _l, _i, _p],
    "t10_relayout_segments": [_i, _i, _p, _p, _l, _p, _p],
    "t10_svd_small_limits": [_p, _p],
    "t10_svd_small": [_i, _l, _p, _p, _p, _p, _p, _p, _p, _p],
    "t10_ggemm_tma_slots": [_p],
    "t10_ggemm_tma_encode": [_l, _p, _p, _p, _p],
    "t10_ggemm_tma": [_p, _p, _l, _p, _l, _p, _p, _l, _p, _p, _p, _i, _i, _i, _p, C.c_uint64, _p, _i, _p, _p],
    "t10_signal_flags": [_i, _p, C.c_uint64, _p],
    "t10_wait_flags": [_i, _p, _p, _l, _p, _p],
    "t10_relayout_runs": [_i, _p, _p, _l, _p, _l, _p, _p],
    "t10_ggemm_allgather": [_i, _p, _p, _i, _i, _p, _p, C.c_uint64, _p, _l, _p, _l, _p, _l, _p, _p, _p],
    "t10_wait_copy": [_i, _p, _p, _l, _p, _i, C.c_uint64, _p, _p],
    "t10_probe_f64_peak": [_i, _i, _p, _p],
    "t10_probe_f64_occupancy": [_i, _i, _p, _p],
}

_lib = None


def load():
    """dlopen the library (no CUDA call is made).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "tor10_b200: %s is missing -- run `python tor10_b200/build.py` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = C.c_char_p if name == "t10_last_error" else C.c_int
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().t10_last_error().decode("utf-8", "replace")
        if rc == 4:
            raise NoValidBlock("UniTensor.__init__", "[ERROR] no vaild block in current Tensor. " + msg)
        raise T10Error(rc, msg)


def require_cuda(dev):
    if not torch.cuda.is_available():
        raise RuntimeError("tor10_b200 needs a CUDA device (sm_100a); there is no CPU path")
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise RuntimeError("tor10_b200 kernels run on CUDA tensors only (got device %s)" % dev)
    return dev


def require_cuda_tensor(x, what):
    if x.device.type != "cuda" or x.dtype not in (torch.float64, torch.float32):
        raise RuntimeError("tor10_b200: %s needs a CUDA float64/float32 tensor (got %s on %s); there is no CPU "
                           "path" % (what, x.dtype, x.device))


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw handle, no Stream object)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class on_device:
    """`with torch.cuda.device(d)` that does nothing (and costs nothing) when d is already current."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.type == "cuda" else None
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch._C._cuda_getDevice()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def i64_args(vals):
    """small int64 host array for an ABI call (cheaper than building a numpy array)"""
    return (C.c_int64 * len(vals))(*vals)


def dtype_code(dt):
    if dt == torch.float64:
        return T10_F64
    if dt == torch.float32:
        return T10_F32
    raise TypeError("tor10_b200 kernels support torch.float64 / torch.float32, got %s" % dt)


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def i64(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)
