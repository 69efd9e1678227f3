"""Tracing front end: a `hecate`-compatible DSL that records CKKS programs as a plain op graph.

Same user-facing surface as the reference's Python DSL (reference: python/hecate/hecate/expr.py:
`Expr` operator overloading 145-204, `Plain` 260-274, `Empty` 276-290, `func` 297-306, `save` 92-110,
`bootstrap` 115-127), but instead of building MLIR through libHecateFrontend.so (tools/frontend.cpp)
it appends nodes to a Python list which `dacapo_b200.compiler` lowers to an HEVM program.

    import dacapo_b200.frontend as hc
    @hc.func("c")
    def f(x):
        return (x * [0.5, 0.25]).rotate(1) + x
    graph = hc.save()          # -> Graph(nodes, consts, outputs)

Node kinds: ("input", i) ("const", k) ("add", a, b) ("sub", a, b) ("mul", a, b) ("neg", a)
("rot", a, offset) ("boot", a).  `rotate(k)` rotates the slot vector LEFT by k (SEAL rotate_vector,
SEAL_HEVM.cpp:273); `sub` is lowered later as negate + add like tools/frontend.cpp:192-198.
"""
from collections.abc import Iterable
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Graph:
    nodes: List[tuple] = field(default_factory=list)
    consts: List[np.ndarray] = field(default_factory=list)
    outputs: List[int] = field(default_factory=list)
    n_inputs: int = 0


_G = Graph()
_FUNCS = []


def _emit(*node):
    _G.nodes.append(tuple(node))
    return Expr(len(_G.nodes) - 1)


class Expr:
    """A traced value (ciphertext or plaintext constant)."""

    def __init__(self, idx):
        self.idx = idx

    def _bin(self, op, other, reverse=False):
        o = resolveType(other)
        a, b = (o.idx, self.idx) if reverse else (self.idx, o.idx)
        return _emit(op, a, b)

    def __add__(self, o):
        return self._bin("add", o)

    def __sub__(self, o):
        return self._bin("sub", o)

    def __mul__(self, o):
        return self._bin("mul", o)

    # the reference binds both __r<op>__ and __i<op>__ to the reversed form (expr.py:158-177)
    def __radd__(self, o):
        return self._bin("add", o, True)

    def __rsub__(self, o):
        return self._bin("sub", o, True)

    def __rmul__(self, o):
        return self._bin("mul", o, True)

    __iadd__ = __radd__
    __isub__ = __rsub__
    __imul__ = __rmul__

    def __neg__(self):
        return _emit("neg", self.idx)

    def rotate(self, offset):
        return _emit("rot", self.idx, int(offset))

    def __copy__(self):
        raise Exception("Copying Hecate object is forbidden")

    __deepcopy__ = __copy__


class Plain(Expr):
    def __init__(self, data, scale=40):
        data = np.array(np.asarray(data).tolist(), dtype=np.float64).ravel()
        _G.consts.append(data)
        _G.nodes.append(("const", len(_G.consts) - 1))
        super().__init__(len(_G.nodes) - 1)


def resolveType(other):
    if isinstance(other, Expr):
        return other
    if isinstance(other, (int, float)):
        return Plain(np.array([other], dtype=np.float64))
    try:
        import torch
        if isinstance(other, torch.Tensor):
            return Plain(torch.flatten(other).tolist())
    except ImportError:
        pass
    return Plain(other)


class Empty:
    """Additive identity placeholder (expr.py:276-290)."""

    def __add__(self, other):
        return resolveType(other)

    __radd__ = __iadd__ = __sub__ = __rsub__ = __isub__ = __add__


LIVE_LIST_BOOTSTRAPS = False


def bootstrap(x):
    """Manual bootstrap hint.  For iterables the reference creates one bootstrap op per element but returns the
    ARGUMENT, so those ops are dead and disappear in the first canonicalisation (expr.py:115-127); mirrored: nothing is
    recorded.  Every `hc.bootstrap(out)` of examples/benchmarks/ResNet.py is of that kind, which is why only the
    `dacapo` pipeline (automatic placement) can compile it.  With LIVE_LIST_BOOTSTRAPS = True the hint is honoured
    element-wise instead -- the hand placement the benchmark's author wrote down, used as the `pars` arm of the
    DaCapo-vs-PARS comparison (BASELINE.json configs[3])."""
    if isinstance(x, Expr):
        return _emit("boot", x.idx)
    if LIVE_LIST_BOOTSTRAPS and isinstance(x, Iterable):
        if isinstance(x, np.ndarray):
            out = np.empty(x.shape, dtype=object)
            for idx in np.ndindex(x.shape):
                out[idx] = bootstrap(x[idx])
            return out
        return type(x)(bootstrap(e) for e in x)
    return x


def func(param):
    def gen(f):
        _FUNCS.append((f, param))
        return f
    return gen


def save(dirs="", cst_dirs=""):
    """Evaluate every registered @func and return the recorded Graph (the reference writes .mlir/.cst here)."""
    global _G
    for f, param in _FUNCS:
        ins = []
        for a in param.split(","):
            if a.strip() != "c":
                raise ValueError("only ciphertext parameters ('c') are supported")
            _G.nodes.append(("input", _G.n_inputs))
            _G.n_inputs += 1
            ins.append(Expr(len(_G.nodes) - 1))
        ret = f(*ins)
        if isinstance(ret, Expr):
            ret = [ret]
        for r in np.ravel(np.asarray(ret, dtype=object)) if not isinstance(ret, list) else ret:
            if isinstance(r, Iterable) and not isinstance(r, Expr):
                for rr in np.ravel(np.asarray(r, dtype=object)):
                    _G.outputs.append(rr.idx)
            else:
                _G.outputs.append(r.idx)
    _FUNCS.clear()
    g, _G = _G, Graph()
    return g


def reset():
    global _G
    _G = Graph()
    _FUNCS.clear()


def evaluate(graph: Graph, inputs, nslots: int):
    """Plaintext semantics of a traced graph on slot vectors (the functional reference of a compiled program):
    inputs and constants are tiled `src[i % len]` over the `nslots` slots, as the runtime's `encrypt` and `preprocess`
    both do through `encode_internal` (SEAL_HEVM.cpp:256-261, 439-445); `rot` rotates left.  Returns one vector per
    graph output."""
    vals = {}
    for i, n in enumerate(graph.nodes):
        k = n[0]
        if k == "input":
            v = np.resize(np.asarray(inputs[n[1]], dtype=np.float64).ravel(), nslots)
        elif k == "const":
            v = np.resize(np.asarray(graph.consts[n[1]], dtype=np.float64).ravel(), nslots)
        elif k == "add":
            v = vals[n[1]] + vals[n[2]]
        elif k == "sub":
            v = vals[n[1]] - vals[n[2]]
        elif k == "mul":
            v = vals[n[1]] * vals[n[2]]
        elif k == "neg":
            v = -vals[n[1]]
        elif k == "rot":
            v = np.roll(vals[n[1]], -n[2])
        elif k == "boot":
            v = vals[n[1]]
        else:
            raise ValueError(k)
        vals[i] = v
    return [vals[o] for o in graph.outputs]
