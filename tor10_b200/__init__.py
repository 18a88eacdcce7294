"""tor10_b200: the block-sparse contraction path of Tor10 on B200 (sm_100a) kernels.

    import tor10_b200 as tor10      # same names as the reference package

Public surface = the reference's (`tor10/__init__.py`): Bond, BD_KET/BD_BRA/BD_REG, Symmetry,
UniTensor, Contract, Save/Load/From_torch, linalg (Matmul, Chain_matmul, Svd_truncate, ExpH, ...),
Network.  The compute path is the C-ABI library built from tor10_b200/csrc (see include/tor10_b200.h);
importing the package loads it and fails loudly if it has not been built.
"""
from . import _lib

_lib.load()  # no CPU fallback: a missing libtor10_b200.so is an import error

from . import Symmetry  # noqa: E402
from .Bond import BD_BRA, BD_KET, BD_REG, Bond, BondType, SymmetryTypes, _fx_GetCommRows  # noqa: E402,F401
from .UniTensor import (Contract, From_torch, Load, Save, UniTensor, _CombineBonds, _Randomize)  # noqa: E402,F401
from .linalg import (Abs, Chain_matmul, Det, ExpH, Hosvd, Inverse, Matmul, Mean, Norm, Otimes, Qdr, Qr, Svd,  # noqa: E402,F401
                     Svd_truncate)
from .Network import Network  # noqa: E402,F401
from . import _plan  # noqa: E402,F401

__version__ = "0.1.0"
