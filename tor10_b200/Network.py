"""Network: contraction-order executor (API of tor10/Network.py).

File format, `Put`, `Launch` and the evaluation order are those of the reference:
  * `.net` text: `Name: in-labels ; out-labels`, one `TOUT:` line, optional `Order:` line (Network.py:73-148);
  * with an `Order`, a shunting-yard evaluator pops the MOST RECENT operand first, i.e. it calls
    Contract(later, earlier) (Network.py:349-351); without one, tensors are folded in the order they
    were `Put` (Network.py:404-415);
  * instance labels are replaced by the net-file labels for the duration of the launch and restored.

Extensions over the reference: tagged / symmetric tensors are accepted (the reference
refuses them, Network.py:301-302), `Network(nwfile)` works (the reference calls a misspelt method,
:68), and every pairwise Contract replays a cached ContractPlan, so a second Launch of the same
network does no block-map work at all.
"""
import copy
import re

import numpy as np

from .UniTensor import Contract, UniTensor


class Network:
    def __init__(self, nwfile=None, delimiter=None):
        self.tensors = None
        self.TOUT = None
        self.instances = None
        self.Order = None
        if nwfile is not None:
            self.Fromfile(nwfile, delimiter)

    # ------------------------------------------------------------------ parsing ------
    @staticmethod
    def _is_matched(expression):
        depth = 0
        for ch in expression:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth < 0:
                    return False
        return depth == 0

    def Fromfile(self, ntwrk_file, delimiter=None):
        with open(ntwrk_file, "r") as f:
            self._parse(f.readlines(), delimiter)

    def Fromstring(self, text, delimiter=None):
        """Same grammar as Fromfile, from an in-memory string."""
        self._parse(text.splitlines(), delimiter)

    def _parse(self, lines, delimiter):
        self.tensors, self.TOUT, self.instances, self.Order = None, None, None, None
        for i, line in enumerate(lines):
            line = line.strip()
            if line == "":
                continue
            parts = line.split(":")
            if len(parts) != 2:
                raise TypeError("Network.fromfile", "[ERROR] The network file have wrong format at line [%d]" % i)
            name = parts[0].strip()
            if name == "Order":
                if self.Order is not None:
                    raise TypeError("Network.fromfile", "[ERROR] The network file have multiple line of [Order]")
                order = parts[1].strip()
                if not self._is_matched(order):
                    raise TypeError("Network.fromfile",
                                    "[ERROR] The parentheses mismatch happend for the [Order] at line [%d]" % i)
                self.Order = order
                continue
            halves = parts[1].split(";")
            if len(halves) != 2:
                if name != "TOUT":
                    raise Exception("Network.fromfile", "[ERROR] syntax error for Network file.")
                shell = [None, None]
            else:
                shell = [np.array(h.strip().split(delimiter), dtype=np.int64) for h in halves]
            if name == "TOUT":
                if self.TOUT is not None:
                    raise TypeError("Network.fromfile", "[ERROR] The network file have multiple line of [TOUT]")
                self.TOUT = shell
            else:
                if self.tensors is None:
                    self.tensors = {}
                if name in self.tensors:
                    raise ValueError("Network.fromfile",
                                     "[ERROR] network file contain duplicate tensor names [%s] at line [%d]" % (name, i))
                self.tensors[name] = shell
        if self.TOUT is None:
            raise TypeError("Network.fromfile", "[ERROR] network file have no TOUT element.")
        if self.tensors is None:
            raise TypeError("Network.fromfile", "[ERROR] network file have no input elements exist.")

    def __repr__(self):
        print("==== Network ====")
        if self.tensors is None:
            print("      Empty      ")
        else:
            for key, val in self.tensors.items():
                status = "o" if (self.instances is not None and key in self.instances) else "x"
                print("[%s] %s : " % (status, key) + " ".join("%d" % i for i in val[0]) + " ; " +
                      " ".join("%d" % i for i in val[1]))
            if self.TOUT[0] is not None:
                print("TOUT : " + " ".join("%d" % i for i in self.TOUT[0]) + " ; " +
                      " ".join("%d" % i for i in self.TOUT[1]))
        print("=================")
        return ""

    __str__ = __repr__

    # ------------------------------------------------------------------ Put / Launch -
    def Put(self, name, tensor):
        if self.tensors is None:
            raise ValueError("Network.put", "[ERROR] Network hasn't been constructed. Construct a Netwrok before put "
                                            "the tensors.")
        if not isinstance(tensor, UniTensor):
            raise TypeError("Network.put", "[ERROR] Network can only accept UniTensor")
        if name not in self.tensors:
            raise ValueError("Network.put", "[ERROR] Network does not contain the tensor with name [%s]" % name)
        if len(tensor.shape) != len(self.tensors[name][0]) + len(self.tensors[name][1]):
            raise TypeError("Network.put", "[ERROR] Trying to put tensor %s that has different rank" % name)
        if self.instances is None:
            self.instances = {}
        self.instances[name] = tensor

    def _net_labels(self, key):
        return np.array(self.tensors[key][0].tolist() + self.tensors[key][1].tolist(), dtype=np.int64)

    def _launch_by_order(self):
        tokens = re.findall(r"[(,)]|\w+", self.Order)
        old_labels = {}
        for key in self.tensors:
            old_labels[key] = copy.copy(self.instances[key].labels)
            self.instances[key].labels = self._net_labels(key)
        try:
            values, operators = [], []

            def reduce_once():
                left = values.pop()      # most recent operand first (Network.py:349-351)
                right = values.pop()
                values.append(Contract(left, right))

            for token in tokens:
                if token == "(":
                    operators.append(token)
                elif token == ")":
                    while operators and operators[-1] != "(":
                        operators.pop()
                        reduce_once()
                    operators.pop()
                elif token == ",":
                    while operators and operators[-1] not in "()":
                        operators.pop()
                        reduce_once()
                    operators.append(token)
                else:
                    if token not in self.instances:
                        raise ValueError("Network.Launch", "[ERROR] Order names an unknown tensor [%s]" % token)
                    values.append(self.instances[token])
            while operators:
                operators.pop()
                reduce_once()
        finally:
            for key in self.tensors:
                self.instances[key].labels = old_labels[key]
        return values[0]

    def Launch(self):
        if self.tensors is None:
            raise Exception("Network", "[ERROR] No in-put tensors for the Network")
        if self.TOUT is None:
            raise Exception("Network", "[ERROR] No TOUT tensor for the Network")
        if self.instances is None:
            raise Exception("Network", "[ERROR] No UniTensor is put in the Network")
        for key in self.tensors.keys():
            if key not in self.instances:
                raise Exception("Network", "[ERROR] The [%s] tensor is not put in the Network" % key)
        out = None
        if self.Order is None:
            for key, value in self.instances.items():
                if out is None:
                    out = copy.deepcopy(value)
                    out.labels = self._net_labels(key)
                else:
                    old = copy.copy(value.labels)
                    value.labels = self._net_labels(key)
                    try:
                        out = Contract(out, value)
                    finally:
                        value.labels = old
        else:
            out = self._launch_by_order()
            if len(self.tensors) == 1:
                out = copy.deepcopy(out)
        if len(out.shape) == 0:
            return out
        per_lbl = self.TOUT[0].tolist() + self.TOUT[1].tolist()
        out.Permute(per_lbl, len(self.TOUT[0]), by_label=True)
        return out
