"""Bond: index descriptor (API of tor10/Bond.py).

Same constructor, attributes (`dim`, `bondType`, `qnums`, `nsym`, `sym_types`) and methods as the
reference.  Differences that matter to the B200 path:
  * `qnums` is an immutable (read-only) int64 array.  Every mutating method (`assign`, `combine`,
    `__mul__`) installs a new array, so a bond has a cheap cached signature that keys the
    block-map / contraction plan caches (tor10_b200/_plan.py).
  * copies share the immutable array (the reference deep-copies every bond into every tensor,
    UniTensor.py:263).
Host-side fusion here (`combine`) is metadata for users of the Bond API; the block map itself is
fused on the device (csrc/blockmap.cu).
"""
import copy

import numpy as np

from . import Symmetry as Symm


def _fx_GetCommRows(A, B):
    """Row-set intersection, ascending lexicographic (Bond.py:13-28)."""
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    dtype = {"names": ["f{}".format(i) for i in range(A.shape[1])], "formats": A.shape[1] * [A.dtype]}
    Cm = np.intersect1d(A.view(dtype), B.view(dtype))
    return Cm.view(A.dtype).reshape(-1, A.shape[1])


class BD_KET:
    pass


class BD_BRA:
    pass


class BD_REG:
    pass


BondType = {BD_BRA: 1, BD_KET: -1, BD_REG: 0}
SymmetryTypes = {"U1": Symm.U1, "Zn": Symm.Zn}
_TYPE_OF = {1: BD_BRA, -1: BD_KET, 0: BD_REG}


def _freeze(q):
    q = np.ascontiguousarray(q, dtype=np.int64)
    q.setflags(write=False)
    return q


class Bond:
    def __init__(self, dim, bondType=BD_REG, qnums=None, sym_types=None):
        self.bondType = bondType
        self.dim = dim
        self.qnums = None
        self.nsym = 0
        self.sym_types = None
        self._sig = None
        self.assign(dim, bondType, qnums, sym_types)

    def assign(self, dim, bondType=BD_REG, qnums=None, sym_types=None):
        """Bond.py:182-284 (validation order and messages follow the reference)."""
        if dim < 1:
            raise Exception("Bond.assign()", "[ERROR] Bond dimension must be greater than 0.")
        if bondType not in BondType:
            raise Exception("Bond.assign()", "[ERROR] bondType can only be BD_BRA, BD_KET or BD_REG")
        if qnums is not None:
            if bondType is BD_REG:
                raise Exception("Bond.assign()", "[ERROR] with qnums, bondType can only be BD_BRA or BD_KET")
            try:
                lens = np.unique([len(qnums[x]) for x in range(len(qnums))]).flatten()
            except TypeError:
                raise TypeError("Bond.assign()", "[ERROR] qnums must be a list of lists (2D list).")
            if len(lens) != 1:
                raise TypeError("Bond.assign()", "[ERROR] Number of symmetries must be the same for each dim.")
            if len(np.shape(qnums)) != 2:
                raise TypeError("Bond.assign()", "[ERROR] qnums must be a list of lists (2D list).")
            if len(qnums) != dim:
                raise ValueError("Bond.assign()", "[ERROR] qnums must have the same elements as dim")
            nsym = int(lens[0])
            q = np.array(qnums).astype(np.int64)
            if sym_types is None:
                syms = np.array([SymmetryTypes["U1"]() for _ in range(nsym)])
            else:
                if nsym != len(sym_types):
                    raise TypeError("Bond.assign()", "[ERROR] Number of symmetry types must match qnums.")
                for s in range(len(sym_types)):
                    if sym_types[s].__class__ not in SymmetryTypes.values():
                        raise TypeError("Bond.assign()", "[ERROR] invalid Symmetry Type.")
                    if not sym_types[s].CheckQnums(q[:, s]):
                        raise TypeError("Bond.assign()", "[ERROR] invalid qnum in Symmetry [%s] @ index: %d" % (
                            str(sym_types[s]), s))
                syms = copy.deepcopy(sym_types)
            # Bond.py:275-276 : rows sorted DESCENDING lexicographically
            y = np.lexsort(q.T[::-1])[::-1]
            self.nsym = nsym
            self.sym_types = syms
            self.qnums = _freeze(q[y, :])
        else:
            if sym_types is not None:
                raise ValueError("Bond.assign()", "[ERROR] the sym_type is assigned but no qnums is passed.")
        self.bondType = bondType
        self.dim = dim
        self._sig = None

    def change(self, new_bondType):
        """Bond.py:289-308"""
        if self.bondType is not new_bondType:
            if new_bondType not in BondType:
                raise TypeError("Bond.change", "[ERROR] the bondtype can only be", BondType)
            if self.qnums is not None and new_bondType == BD_REG:
                raise TypeError("Bond.change", "[ERROR] cannot change a bond with symmetry to BD_REG type")
            self.bondType = new_bondType
            self._sig = None

    def _combine_one(self, bd):
        self.dim *= bd.dim
        if (self.qnums is None) != (bd.qnums is None):
            raise TypeError("Bond.combine", "[ERROR] Trying to combine symmetric and non-symmetric bonds.")
        if self.qnums is not None:
            if self.nsym != len(bd.qnums[0]):
                raise TypeError("Bond.combine", "[ERROR] Trying to combine bonds with different number of symmetries.")
            for s in range(self.nsym):
                if self.sym_types[s] != bd.sym_types[s]:
                    raise TypeError("Bond.combine", "[ERROR] Tryping to combine bonds with different symmetries.")
            # row-major outer fusion, self index major (Bond.py:367-378)
            A = self.qnums.reshape(len(self.qnums), 1, self.nsym)
            B = bd.qnums.reshape(1, len(bd.qnums), self.nsym)
            cols = [self.sym_types[s].CombineRule(A[:, :, s], B[:, :, s]).reshape(-1) for s in range(self.nsym)]
            self.qnums = _freeze(np.stack(cols, axis=1))
        self._sig = None

    def combine(self, bds, new_type=None):
        """Bond.py:310-421"""
        if isinstance(bds, self.__class__):
            self._combine_one(bds)
        else:
            for i in range(len(bds)):
                if not isinstance(bds[i], self.__class__):
                    raise TypeError("Bond.combine(bds)", "bds[%d] is not Bond class" % i)
                self._combine_one(bds[i])
        if new_type is not None:
            if new_type not in BondType:
                raise Exception("Bond.combine(bds,new_type)", "[ERROR] new_type can only be", BondType)
            self.change(new_type)

    def GetUniqueQnums(self, return_degeneracy=False):
        """Bond.py:423-447 (ascending-lex unique rows).  `return_degeneracy=True` returns the
        per-sector multiplicities; the reference's version of that branch compares against the
        whole unique table (Bond.py:443) and is unusable, so the intended result is returned."""
        if self.qnums is None:
            raise TypeError("Bond.GetUniqueQnums", "[ERROR] cannot get qnums from a non-sym bond.")
        if return_degeneracy:
            return np.unique(self.qnums, axis=0, return_counts=True)
        return np.unique(self.qnums, axis=0)

    def GetDegeneracy(self, *qnums):
        if self.qnums is None:
            raise TypeError("Bond.GetDegeneracy", "[ERROR] cannot get degenerate from a non-sym bond.")
        return int((self.qnums == np.array(qnums).astype(np.int64)).all(axis=1).sum())

    def __mul__(self, val):
        """In-place sign flip, no modular reduction for Zn (Bond.py:469-477)."""
        if (val != -1) and (val != 1):
            raise Exception("[ERROR] val can only be +1 or -1")
        self.qnums = _freeze(self.qnums * int(val))
        self._sig = None
        return self

    # ---- copies share the immutable qnums --------------------------------------
    def _clone(self):
        b = Bond.__new__(Bond)
        b.dim, b.bondType, b.qnums, b.nsym = self.dim, self.bondType, self.qnums, self.nsym
        b.sym_types = None if self.sym_types is None else copy.copy(self.sym_types)
        b._sig = self._sig
        return b

    def __copy__(self):
        return self._clone()

    def __deepcopy__(self, memo):
        return self._clone()

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_sig"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        if self.qnums is not None:
            self.qnums = _freeze(self.qnums)
        self._sig = None

    def signature(self):
        """Hashable identity of (dim, tag, charges, symmetries); cached."""
        if self._sig is None:
            if self.qnums is None:
                self._sig = (int(self.dim), BondType[self.bondType], None, None)
            else:
                self._sig = (int(self.dim), BondType[self.bondType], hash(self.qnums.tobytes()),
                             tuple(s._code() for s in self.sym_types))
        return self._sig

    def __print(self):
        out = "Dim = %d |\n" % self.dim
        out += {BD_REG: "REG     :", BD_BRA: "BRA     :", BD_KET: "KET     :"}[self.bondType]
        if self.qnums is not None:
            for n in range(self.nsym):
                out += " %s:: " % (str(self.sym_types[n]))
                out += "".join(" %+d" % q for q in self.qnums[:, n])
                out += "\n         "
        else:
            out += "\n"
        print(out, end="")

    def __str__(self):
        self.__print()
        return ""

    __repr__ = __str__

    def __eq__(self, rhs):
        """Bond.py:509-546"""
        if not isinstance(rhs, self.__class__):
            raise ValueError("Bond.__eq__", "[ERROR] invalid comparison between Bond object and other type class.")
        if not ((self.dim == rhs.dim) and (self.bondType == rhs.bondType)):
            return False
        if (self.qnums is None) != (rhs.qnums is None):
            return False
        if self.qnums is not None:
            if self.qnums.shape != rhs.qnums.shape or not (self.qnums == rhs.qnums).all():
                return False
            return all(self.sym_types[s] == rhs.sym_types[s] for s in range(self.nsym))
        return True

    __hash__ = None
