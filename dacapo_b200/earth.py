"""Python restatement of the reference compiler's `pars` pipeline over a traced op graph:

    Earth typing rules  ->  ProactiveRescaling (PARS)  ->  EarlyModswitch  ->  canonicalise / CSE / DCE  ->
    Earth-to-CKKS level mapping  ->  UpscaleToMulcp  ->  ReuseBuffer  ->  EmitHEVM

`hecate-opt` is MLIR and cannot be built here, so every step is restated from its source (cited per function) on a plain
SSA list.  The values follow the Earth dialect: a value is cipher or plain and carries (scale bits, level), level
counting CONSUMED levels from 0 (EarthOps.td:35-67); the CKKS level of the emitted program is init_level - level
(EarthToCKKS.cpp:155-167).  This is the `pars` arm of BASELINE.json configs[3]; `dacapo_b200.dacapo` adds the DaCapo
bootstrap planner on top of it.  (`dacapo_b200.compiler` is the older, deliberately simpler lowering; it is kept because
the committed ResNet fixtures were produced by it.)
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import hevm_asm as asm
from .frontend import Graph


class PassFailed(Exception):
    """A type-inference rule rejected an op (what makes an MLIR pass `failed()`), e.g. a multiply whose operand has no
    modulus left (EarthDialect.cpp MulOp::inferReturnTypes) -- the signal CandidateSelection relies on."""


@dataclass
class Params:
    """The scalars of a cost profile (EarthDialect.cpp:143-157) + the pass options (optimizer.cpp:206-234)."""
    waterline: int = 40
    output_val: int = 10
    rescaling_factor: int = 60
    level_upper: int = 13            # levelUpperBound
    level_lower: int = 2             # levelLowerBound
    boot_upper: int = 13             # bootstrapLevelUpperBound
    boot_lower: int = 2              # bootstrapLevelLowerBound
    poly_degree: int = 1 << 15
    latency: Dict[str, List[int]] = field(default_factory=dict)   # "earth.<op>_single|_double" -> list indexed by cipher level

    @staticmethod
    def from_profile(prof: dict, waterline=40, output_val=10, exact=False):
        tab = prof.get("latencyTableExact" if exact and "latencyTableExact" in prof else "latencyTable", {})
        # the loader prepends a 0 so that index = cipher level, and pads short lists with their last value (EarthDialect.cpp:158-166)
        top = int(prof["levelUpperBound"])
        lat = {}
        for k, v in tab.items():
            v = [0] + list(v)
            v += [v[-1]] * (top + 2 - len(v))
            lat[k] = v
        return Params(waterline, output_val, int(prof["rescalingFactor"]), top, int(prof["levelLowerBound"]),
                      int(prof["bootstrapLevelUpperBound"]), int(prof["bootstrapLevelLowerBound"]), int(prof["polynomialDegree"]), lat)


class V:
    """One SSA value = one Earth op.  kind: arg const add mul neg rot boot rescale modswitch upscale."""
    __slots__ = ("kind", "ins", "attr", "cipher", "scale", "level", "opid", "tag")

    def __init__(self, kind, ins=(), attr=None, cipher=True, scale=0, level=0):
        self.kind, self.ins, self.attr = kind, list(ins), attr
        self.cipher, self.scale, self.level = cipher, scale, level
        self.opid, self.tag = -1, None

    def __repr__(self):
        return f"%{self.kind}{'' if self.attr is None else '[' + str(self.attr) + ']'}<{'ci' if self.cipher else 'pl'} {self.scale}*{self.level}>"


@dataclass
class Func:
    ops: List[V]
    rets: List[V]
    consts: List[np.ndarray]
    init_level: int = 0

    def clone(self):
        m = {}
        ops = []
        for o in self.ops:
            n = V(o.kind, [m[i] for i in o.ins], o.attr, o.cipher, o.scale, o.level)
            n.opid, n.tag = o.opid, o.tag
            m[o] = n
            ops.append(n)
        return Func(ops, [m[r] for r in self.rets], self.consts, self.init_level), m


def naf(value: int) -> List[int]:
    """Support.h:10-27 (identical to SEAL's)."""
    res, sign, value, i = [], value < 0, abs(value), 0
    while value:
        zi = 2 - (value & 3) if value & 1 else 0
        value = (value - zi) >> 1
        if zi:
            res.append((-zi if sign else zi) * (1 << i))
        i += 1
    return res


# ------------------------------------------------------------------------------------------------ front end -> Earth
def from_graph(g: Graph, keep_bootstraps=True) -> Func:
    """frontend.Graph -> Earth ops.  `sub` becomes negate + add (tools/frontend.cpp:192-198); plain (op) plain is folded;
    x + 0, x * 1, x * (-1) canonicalise like EarthCanonicalizer.td:19-24; every use of a constant gets its own constant op
    (PrivatizeConstant), because scale management types a constant per use."""
    consts: List[np.ndarray] = []
    cid_of = {}

    def cid(arr):
        arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel())
        key = (arr.size, arr.tobytes())
        if key not in cid_of:
            cid_of[key] = len(consts)
            consts.append(arr)
        return cid_of[key]

    ops: List[V] = []
    val = {}      # graph node -> V (cipher) | ("c", ndarray) for plain values
    cse = {}

    def emit(kind, ins, attr=None):
        key = (kind, tuple(id(i) if isinstance(i, V) else ("c", i) for i in ins), attr)
        if key in cse:
            return cse[key]
        real = []
        for i in ins:
            if isinstance(i, V):
                real.append(i)
            else:  # a private constant op for this use
                c = V("const", (), i, cipher=False)
                ops.append(c)
                real.append(c)
        v = V(kind, real, attr)
        ops.append(v)
        cse[key] = v
        return v

    def bc(a, b):
        n = max(a.size, b.size)
        if n % a.size or n % b.size:
            n = int(np.lcm(a.size, b.size))
        return np.resize(a, n), np.resize(b, n)

    for i, n in enumerate(g.nodes):
        k = n[0]
        if k == "input":
            v = V("arg", (), n[1])
            ops.append(v)
            val[i] = v
        elif k == "const":
            val[i] = ("c", np.asarray(g.consts[n[1]], dtype=np.float64).ravel())
        elif k in ("add", "sub", "mul"):
            a, b = val[n[1]], val[n[2]]
            pa, pb = not isinstance(a, V), not isinstance(b, V)
            if pa and pb:
                x, y = bc(a[1], b[1])
                val[i] = ("c", x + y if k == "add" else x - y if k == "sub" else x * y)
                continue
            if k == "sub":  # a - b = a + negate(b)
                b = ("c", -b[1]) if pb else emit("neg", [b])
                pb = not isinstance(b, V)
                k = "add"
            if pa:           # commutative: constant on the right
                a, b, pa, pb = b, a, pb, pa
            if pb:
                c = b[1]
                if k == "add" and not c.any():
                    val[i] = a
                    continue
                if k == "mul" and np.all(c == 1.0):
                    val[i] = a
                    continue
                if k == "mul" and np.all(c == -1.0):
                    val[i] = emit("neg", [a])
                    continue
                val[i] = emit(k, [a, cid(c)])
            else:
                val[i] = emit(k, [a, b])
        elif k == "neg":
            a = val[n[1]]
            val[i] = ("c", -a[1]) if not isinstance(a, V) else emit("neg", [a])
        elif k == "rot":
            a = val[n[1]]
            val[i] = ("c", np.roll(a[1], -n[2])) if not isinstance(a, V) else (a if n[2] == 0 else emit("rot", [a], int(n[2])))
        elif k == "boot":
            a = val[n[1]]
            val[i] = emit("boot", [a], 0) if (keep_bootstraps and isinstance(a, V)) else a
        else:
            raise ValueError(k)
    rets = [val[o] for o in g.outputs]
    if any(not isinstance(r, V) for r in rets):
        raise ValueError("a program output is a plaintext constant")
    f = Func(ops, rets, consts)
    dce(f)
    return f


def dce(f: Func):
    live = set()
    stack = list(f.rets)
    while stack:
        v = stack.pop()
        if id(v) in live:
            continue
        live.add(id(v))
        stack.extend(v.ins)
    f.ops = [o for o in f.ops if id(o) in live or o.kind == "arg"]


def users(f: Func):
    u = {id(o): [] for o in f.ops}
    for o in f.ops:
        for i in o.ins:
            u[id(i)].append(o)
    return u


# ------------------------------------------------------------------------------------------------ type inference
def infer(op: V, P: Params):
    """inferReturnTypes of every Earth op (EarthDialect.cpp:182-340)."""
    Rf = P.rescaling_factor
    k = op.kind
    if k in ("arg", "const"):
        return
    a = op.ins[0]
    if k in ("add", "mul"):
        b = op.ins[1]
        if a.level != b.level:
            raise PassFailed(f"{k}: level mismatch {a} {b}")
        if k == "add":
            if a.scale != b.scale:
                raise PassFailed(f"add: scale mismatch {a} {b}")
            op.scale, op.level = a.scale, a.level
        else:
            if P.boot_upper * Rf < a.level * Rf + a.scale:
                raise PassFailed(f"mul: no modulus left {a}")
            op.scale, op.level = a.scale + b.scale, a.level
        op.cipher = True
    elif k in ("neg", "rot"):
        op.cipher, op.scale, op.level = a.cipher, a.scale, a.level
    elif k == "rescale":
        op.cipher, op.scale, op.level = a.cipher, a.scale - Rf, a.level + 1
    elif k == "modswitch":
        if op.attr < 0:
            raise PassFailed("modswitch: negative downFactor")
        op.cipher, op.scale, op.level = a.cipher, a.scale, a.level + op.attr
    elif k == "upscale":
        if op.attr < 0:
            raise PassFailed("upscale: negative upFactor")
        op.cipher, op.scale, op.level = a.cipher, a.scale + op.attr, a.level
    elif k == "boot":
        if a.level > P.boot_upper - P.boot_lower:
            raise PassFailed(f"bootstrap: operand too deep {a}")
        op.cipher, op.scale, op.level = True, a.scale, op.attr
    else:
        raise ValueError(k)


def _cdiv(a: int, b: int) -> int:
    """C++ integer division (truncates toward zero); a negative result becomes a negative downFactor and fails inference."""
    return a // b if a >= 0 else -((-a) // b)


# ------------------------------------------------------------------------------------------------ PARS
class _Builder:
    def __init__(self, P):
        self.P, self.out = P, []

    def mk(self, kind, a, attr=None):
        v = V(kind, [a], attr)
        infer(v, self.P)
        self.out.append(v)
        return v


def _pars_operands(op: V, ins: List[V], b: _Builder):
    """processOperandsPARS / processOperandsEVA of AddOp, MulOp, BootstrapOp (EarthOps.td:232-255, 346-402, 455-502)."""
    P = b.P
    W, Rf = P.waterline, P.rescaling_factor
    k = op.kind
    if k == "boot":  # BootstrapOp::processOperandsEVA: bring the operand to scale = Rf
        x = ins[0]
        if x.scale + Rf * x.level < (P.boot_upper + 1) * Rf:
            if x.scale < Rf:
                x = b.mk("upscale", x, Rf - x.scale)
            elif x.scale > Rf:
                over = (x.scale - 1) // Rf
                x = b.mk("upscale", x, Rf * (over + 1) - x.scale)
                for _ in range(over):
                    x = b.mk("rescale", x)
        return [x]
    if k not in ("add", "mul"):
        return ins
    x = list(ins)
    both = x[0].cipher and x[1].cipher
    if not both:
        lo = 0 if x[1].cipher else 1
        hi = 1 - lo
        if k == "add":  # the plaintext takes the ciphertext's scale and level, nothing else happens
            x[lo].cipher, x[lo].scale, x[lo].level = False, x[hi].scale, x[hi].level
            return x
        x[lo].cipher, x[lo].scale, x[lo].level = False, W, x[hi].level  # mul: plaintext at the waterline; falls through
    for i in range(2):
        if x[i].scale >= W + Rf:
            x[i] = b.mk("rescale", x[i])
    if x[0].level != x[1].level:
        lo = 0 if x[0].level < x[1].level else 1
        if x[lo].scale != W:
            x[lo] = b.mk("rescale", b.mk("upscale", x[lo], W + Rf - x[lo].scale))
    # processOperandsEVA
    if not both:
        lo = 0 if x[1].cipher else 1
        hi = 1 - lo
        x[lo].scale, x[lo].level = (x[hi].scale if k == "add" else W), x[hi].level
        if k == "add":
            return x
    else:
        if k == "add" and x[0].scale != x[1].scale:
            lo = 0 if x[0].scale < x[1].scale else 1
            x[lo] = b.mk("upscale", x[lo], x[1 - lo].scale - x[lo].scale)
        if x[0].level != x[1].level:
            lo = 0 if x[0].level < x[1].level else 1
            x[lo] = b.mk("modswitch", x[lo], x[1 - lo].level - x[lo].level)
    if k == "mul" and x[0].scale + x[1].scale > 2 * W + Rf:
        x[0] = b.mk("rescale", b.mk("upscale", x[0], W + Rf - x[0].scale))
        if x[0].level != x[1].level:
            x[1] = b.mk("rescale", b.mk("upscale", x[1], W + Rf - x[1].scale))
    return x


def pars(f: Func, P: Params, arg_types=None, mid_segment=False, ret_bypass=None, stop=None) -> Func:
    """ProactiveRescaling.cpp:32-52: refineInputValues, one forward walk (operands, type, results), refineReturnValues
    (Common.cpp:8-103).  `arg_types`: [(scale, level)] per argument (segment_inputType); `stop(op)` may end the walk
    early (BypassDetection / CoverageRecorder) by returning True."""
    W, Rf = P.waterline, P.rescaling_factor
    b = _Builder(P)
    sub = {}
    nargs = 0
    for op in f.ops:
        ins = [sub.get(id(i), i) for i in op.ins]
        if op.kind == "arg":
            op.cipher = True
            op.scale, op.level = arg_types[nargs] if arg_types else (W, 0)
            nargs += 1
            b.out.append(op)
            continue
        if op.kind == "const":
            b.out.append(op)
            continue
        ins = _pars_operands(op, ins, b)
        op.ins = ins
        infer(op, P)
        b.out.append(op)
        if op.kind == "mul":  # MulOp::processResultsEVA: rescale while a whole factor is spare
            cur = op
            while cur.scale >= W + Rf:
                cur = b.mk("rescale", cur)
            if cur is not op:
                sub[id(op)] = cur
        if stop is not None and stop(op):
            f.ops = b.out
            return f
    rets = [sub.get(id(r), r) for r in f.rets]
    # refineReturnValues: results (and bootstrap operands) drop to the lowest level that still holds them
    init_level = P.boot_upper if P.boot_upper >= 0 else P.level_upper

    def refine(vals, out_val, min_level, bypass=None):
        max_req = P.boot_upper - min_level
        if max_req < 0:
            max_req = P.level_upper - min_level
        res = []
        for i, v in enumerate(vals):
            if bypass is not None and bypass[i]:
                res.append(v)
                continue
            diff = _cdiv(max_req * Rf - (v.level * Rf + v.scale + out_val), Rf)
            res.append(b.mk("modswitch", v, diff) if diff else v)
        return res

    if mid_segment:
        rets = refine(rets, 0, P.boot_lower - 1, ret_bypass)
    else:
        rets = refine(rets, P.output_val, 0)
    f.ops, f.rets, f.init_level = b.out, rets, init_level
    # bootstrap operands: refineLevel(bop, waterline, 0, bootstrapLevelLowerBound - 1); the new modswitch goes before the bootstrap
    out2 = []
    for op in f.ops:
        if op.kind == "boot":
            x = op.ins[0]
            max_req = P.boot_upper - (P.boot_lower - 1)
            diff = _cdiv(max_req * Rf - (x.level * Rf + x.scale), Rf)
            if diff:
                m = V("modswitch", [x], diff)
                infer(m, P)
                out2.append(m)
                op.ins = [m]
                infer(op, P)
        out2.append(op)
    f.ops = out2
    return f


# ------------------------------------------------------------------------------------------------ EarlyModswitch + canonicalisation
def early_modswitch(f: Func, P: Params):
    """EarlyModswitch.cpp:36-103: walking backwards, a modswitch that EVERY user of a value applies is moved above the
    value's defining op (onto its operands), so that op runs at the lower level."""
    us = users(f)
    ret_ids = {id(r) for r in f.rets}
    new_before = {}
    for op in reversed(list(f.ops)):
        if op.kind in ("boot", "arg"):
            continue
        uu = us[id(op)]
        if not uu or id(op) in ret_ids:
            continue
        mn = None
        for u in uu:
            if u.kind == "modswitch" and u.ins[0] is op:
                mn = u.attr if mn is None else min(mn, u.attr)
            else:
                mn = 0
        if not mn:
            continue
        if op.kind == "const":
            op.level += mn
        elif op.kind == "modswitch":
            op.attr += mn
            op.level += mn
        else:
            pre, made = [], {}
            for i, x in enumerate(op.ins):
                if x.kind == "const":
                    x.level += mn
                    continue
                if id(x) in made:  # x * x: one modswitch feeds both operands
                    op.ins[i] = made[id(x)]
                    us[id(made[id(x)])].append(op)
                    continue
                m = V("modswitch", [x], mn)
                infer(m, P)
                pre.append(m)
                made[id(x)] = m
                op.ins[i] = m
                seen_op = False
                nu = []
                for t in us[id(x)]:  # the (possibly repeated) uses by `op` collapse into one use by `m`
                    if t is op:
                        if not seen_op:
                            nu.append(m)
                            seen_op = True
                    else:
                        nu.append(t)
                us[id(x)] = nu
                us[id(m)] = [op]
            new_before[id(op)] = pre
            op.level += mn
        for u in uu:
            if u.kind == "modswitch":
                u.attr -= mn
    out = []
    for op in f.ops:
        out.extend(new_before.get(id(op), []))
        out.append(op)
    f.ops = out
    # the walk visits newly created modswitches too (they sit before their op): emulate by iterating to a fixed point
    return f


def canonicalize(f: Func, P: Params):
    """ZeroUpscale / ZeroModswitch / UpscaleUpscale / ModswitchModswitch / RescaleUpscale (EarthCanonicalizer.td:19-58),
    constants absorb upscale / modswitch, then CSE and DCE."""
    Rf = P.rescaling_factor
    changed = True
    while changed:
        changed = False
        us = users(f)
        sub = {}
        for op in f.ops:
            op.ins = [sub.get(id(i), i) for i in op.ins]
            if op.kind in ("upscale", "modswitch"):
                x = op.ins[0]
                if op.attr == 0:
                    sub[id(op)] = x
                    changed = True
                elif x.kind == op.kind:
                    op.ins, op.attr = [x.ins[0]], op.attr + x.attr
                    changed = True
                elif x.kind == "const" and len(us[id(x)]) == 1:
                    if op.kind == "upscale":
                        x.scale += op.attr
                    else:
                        x.level += op.attr
                    sub[id(op)] = x
                    changed = True
                elif op.kind == "upscale" and x.kind == "rescale" and len(us[id(x)]) == 1:
                    # UpscaleRescalePattern: upscale(rescale(x), up) -> rescale(upscale(x, up))
                    u = V("upscale", [x.ins[0]], op.attr)
                    infer(u, P)
                    f.ops.insert(f.ops.index(op), u)
                    op.kind, op.attr, op.ins = "rescale", None, [u]
                    infer(op, P)
                    changed = True
                    break
            elif op.kind == "rescale":
                x = op.ins[0]
                if x.kind == "upscale" and x.attr - Rf >= 0:  # rescale(upscale(x, up)) -> modswitch(upscale(x, up - Rf), 1)
                    op.kind, op.attr = "modswitch", 1
                    if x.attr - Rf == 0:
                        op.ins = [x.ins[0]]
                    else:
                        u = V("upscale", [x.ins[0]], x.attr - Rf)
                        infer(u, P)
                        f.ops.insert(f.ops.index(op), u)
                        op.ins = [u]
                    infer(op, P)
                    changed = True
                    break
        f.rets = [sub.get(id(r), r) for r in f.rets]
        for op in f.ops:
            op.ins = [sub.get(id(i), i) for i in op.ins]
        dce(f)
    # CSE (constants are keyed by their final type)
    seen, sub, out = {}, {}, []
    for op in f.ops:
        op.ins = [sub.get(id(i), i) for i in op.ins]
        if op.kind == "arg":
            out.append(op)
            continue
        key = (op.kind, tuple(id(i) for i in op.ins), op.attr, op.cipher, op.scale, op.level)
        if key in seen:
            sub[id(op)] = seen[key]
        else:
            seen[key] = op
            out.append(op)
    f.ops = out
    f.rets = [sub.get(id(r), r) for r in f.rets]
    dce(f)
    return f


def verify(f: Func, P: Params):
    """Re-run type inference over the final function: every rule must hold and no level may exceed the budget."""
    for op in f.ops:
        if op.kind in ("arg", "const"):
            continue
        s, l, c = op.scale, op.level, op.cipher
        infer(op, P)
        if (s, l) != (op.scale, op.level):
            raise PassFailed(f"stale type on {op}: had {s}*{l}")
        if op.cipher and op.level > f.init_level - 1 and op.kind != "boot":
            raise PassFailed(f"level budget exceeded: {op} (init_level {f.init_level})")
    return True


# ------------------------------------------------------------------------------------------------ latency
_LAT_NAME = {"add": "earth.add", "mul": "earth.mul", "neg": "earth.negate", "rot": "earth.rotate", "boot": "earth.bootstrap",
             "rescale": "earth.rescale", "modswitch": "earth.modswitch", "upscale": "earth.upscale", "const": "earth.constant"}


def op_latency(op: V, P: Params, init_level: int) -> float:
    """HEProfInterface::getLatency x getNum (LatencyEstimator.cpp:27-37, EarthOps.td:79-89, 195-209, 280-292)."""
    if op.kind == "arg":
        return 0.0
    single = not (op.kind in ("add", "mul") and op.ins[0].cipher and op.ins[1].cipher)
    tab = P.latency.get(_LAT_NAME[op.kind] + ("_single" if single else "_double"))
    if not tab:
        return 0.0
    lvl = init_level - op.level

    def at(l):
        return tab[max(0, min(l, len(tab) - 1))]

    if op.kind == "modswitch":
        return float(sum(at(lvl - i) for i in range(op.attr)))  # getLatency of ModswitchOp: one table entry per dropped level
    if op.kind == "rot":
        half = P.poly_degree // 2
        n = sum(1 for t in naf(op.attr % half) if t % half)
        return float(at(lvl) * n)
    return float(at(lvl))


def latency(f: Func, P: Params) -> float:
    return sum(op_latency(o, P, f.init_level) for o in f.ops)


# ------------------------------------------------------------------------------------------------ pipelines + emission
def run_pars_pipeline(f: Func, P: Params) -> Func:
    """optimizer.cpp:437-480 up to the Earth level: ProactiveRescaling, EarlyModswitch, CSE, canonicalise."""
    pars(f, P)
    early_modswitch(f, P)
    canonicalize(f, P)
    verify(f, P)
    return f


def op_counts(f: Func) -> Dict[str, int]:
    c = {}
    for o in f.ops:
        if o.kind in ("arg", "const"):
            continue
        k = o.kind
        if k in ("add", "mul"):
            k += "_cc" if (o.ins[0].cipher and o.ins[1].cipher) else "_cp"
        c[k] = c.get(k, 0) + 1
    c["rotate_keyswitch_steps"] = sum(len(naf(o.attr)) for o in f.ops if o.kind == "rot")
    return c


def emit_hevm(f: Func, P: Params) -> asm.Program:
    """EarthToCKKS (level = init_level - earth level, EarthToCKKS.cpp:155-167; bootstrap target = levelUpperBound - level,
    318-324) + UpscaleToMulcp (Encode(all-ones, 2^upFactor) + MulCP, UpscaleToMulcp.cpp:52-72) + ReuseBuffer
    (liveness-based register reuse, ReuseBuffer.cpp:27-55) + EmitHEVM (EmitHEVM.cpp:31-119)."""
    init = f.init_level
    p = asm.Program(init_level=init)
    last_use = {}
    for idx, o in enumerate(f.ops):
        for i in o.ins:
            last_use[id(i)] = idx
    for r in f.rets:
        last_use[id(r)] = len(f.ops)
    reg, free, pts, pool = {}, [], {}, {}
    args = [o for o in f.ops if o.kind == "arg"]
    for a in sorted(args, key=lambda o: o.attr):
        reg[id(a)] = p.arg(a.scale, init - a.level)

    def plain_reg(cid, level, scale):
        key = (cid, level, scale)
        if key not in pts:
            r = p.new_pt()
            if cid >= 0 and cid not in pool:
                pool[cid] = p.const(f.consts[cid])
            p.encode(r, -1 if cid < 0 else pool[cid], level, scale)
            pts[key] = r
        return pts[key]

    opc = {"neg": asm.NEGATE, "rescale": asm.RESCALE, "modswitch": asm.MODSWITCH, "boot": asm.BOOTSTRAP}
    arg_ids = {id(a) for a in args}
    for idx, o in enumerate(f.ops):
        if o.kind in ("arg", "const"):
            continue
        srcs = [i for i in o.ins if i.cipher]
        for s in {id(s) for s in srcs}:
            if last_use.get(s) == idx and s not in arg_ids:
                free.append(reg[s])
        r = free.pop() if free else p.new_ct()
        reg[id(o)] = r
        lvl = init - o.ins[0].level if o.ins else init
        if o.kind in ("add", "mul"):
            a, b = o.ins
            if a.cipher and b.cipher:
                p.emit(asm.ADDCC if o.kind == "add" else asm.MULCC, r, reg[id(a)], reg[id(b)])
            else:
                ct, pl = (a, b) if a.cipher else (b, a)
                p.emit(asm.ADDCP if o.kind == "add" else asm.MULCP, r, reg[id(ct)], plain_reg(pl.attr, init - pl.level, pl.scale))
        elif o.kind == "rot":
            p.rotate(r, reg[id(o.ins[0])], o.attr)
        elif o.kind == "upscale":
            p.emit(asm.MULCP, r, reg[id(o.ins[0])], plain_reg(-1, lvl, o.attr))
        elif o.kind == "boot":
            p.emit(asm.BOOTSTRAP, r, reg[id(o.ins[0])], P.level_upper - o.level)
        elif o.kind == "modswitch":
            p.emit(asm.MODSWITCH, r, reg[id(o.ins[0])], o.attr)
        else:
            p.emit(opc[o.kind], r, reg[id(o.ins[0])], 0)
    for v in f.rets:
        p.result(reg[id(v)], v.scale, init - v.level)
    return p


def compile_pars(graph: Graph, P: Optional[Params] = None):
    """`hopts pars <waterline>`: keeps the program's manual bootstraps (optimizer.cpp:437-480)."""
    P = P or Params()
    f = from_graph(graph, keep_bootstraps=True)
    run_pars_pipeline(f, P)
    return emit_hevm(f, P), f
