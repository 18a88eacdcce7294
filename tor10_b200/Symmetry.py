"""Fusion rules of the abelian symmetries (API of tor10/Symmetry.py).

Host side only: the device restatement of the same rules lives in csrc/blockmap.cu
(`fuse_row`).  U1: a+b (Symmetry.py:33-34).  Zn: (a+b) % n with numpy's non-negative remainder
(Symmetry.py:83-84); valid charges are 0 <= q < n (Symmetry.py:101-105).
"""
import numpy as np


class U1:
    def __init__(self):
        pass

    def CombineRule(self, A, B):
        return A + B

    def __repr__(self):
        return "U1"

    __str__ = __repr__

    def __eq__(self, rhs):
        return self.__class__ == rhs.__class__

    def __hash__(self):
        return hash("U1")

    def CheckQnums(self, qlist):
        return True

    def _code(self):
        return (0, 0)


class Zn:
    def __init__(self, n):
        if n < 2:
            raise ValueError("Symmetry.Zn", "[ERROR] discrete symmetry Zn must have n >= 2.")
        self.n = int(n)

    def CombineRule(self, A, B):
        return (A + B) % self.n

    def __repr__(self):
        return "Z%d" % self.n

    __str__ = __repr__

    def __eq__(self, rhs):
        return self.__class__ == rhs.__class__ and self.n == rhs.n

    def __hash__(self):
        return hash(("Zn", self.n))

    def CheckQnums(self, qlist):
        q = np.asarray(qlist)
        return bool(((q >= 0) & (q < self.n)).all())

    def _code(self):
        return (1, self.n)
