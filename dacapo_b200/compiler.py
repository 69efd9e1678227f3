"""A small CKKS compiler: traced op graph (dacapo_b200.frontend.Graph) -> HEVM program (.hevm/.cst).

What the reference does with `hecate-opt` (MLIR; tools/optimizer.cpp:236-480) and what is restated here,
in plain Python and deliberately simpler:

  * constant folding / canonicalisation (EarthCanonicalizer.td:19-46: x+0, x*1; plain (op) plain)
  * `sub` -> negate + add (tools/frontend.cpp:192-198)
  * scale management in the spirit of PARS (ProactiveRescaling.cpp:32-52, EarthOps.td:378-402): a
    waterline W (default 2^40) and the rescaling factor R = 60 bits; products are left at high scale and
    rescaled lazily (one rescale after a sum of products, not one per product); additions align scales
    (rescale the larger operand, or upscale the smaller one: Encode(ones, 2^k) + MulCP as in
    UpscaleToMulcp.cpp:52-72) and levels (modswitch, EarlyModswitch.cpp)
  * bootstrap placement: greedy "as late as possible" -- a ciphertext is bootstrapped only when the next
    operation would not fit its modulus.  (DaCapo's planner, DaCapoPlanner.cpp:39-219, searches placements
    with the measured cost table instead; not restated.)
  * Earth -> CKKS level convention (level = number of data limbs, EarthToCKKS.cpp:159-166), liveness based
    register reuse (ReuseBuffer.cpp:27-55) and the HEVM container (EmitHEVM.cpp:31-119)

The output runs on any HEVM runtime with the reference ABI (libB200_HEVM.so, the CPU oracle, libSEAL_HEVM.so).
"""
import math
from dataclasses import dataclass

import numpy as np

from . import hevm_asm as asm
from .frontend import Graph


def _is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def prime_chain(logN, bits, count):
    """SEAL CoeffModulus::Create order: ascending, the last one is the special prime."""
    step = 2 << logN
    v = ((1 << bits) - 1) // step * step + 1
    out = []
    while len(out) < count:
        if _is_prime(v):
            out.append(v)
        v -= step
    return out[::-1]


@dataclass
class Options:
    logN: int = 15
    num_primes: int = 14
    waterline: int = 40
    rescale_bits: int = 60
    margin_bits: int = 10      # head room for |value| and noise: scale_bits + margin <= 60*level - 1
    boot_target: int = 0       # level after a greedy (in-segment) bootstrap (0 = top level)
    plain_bits: int = 40       # default scale of a plaintext factor (encoding error ~ sqrt(N/12) / 2^bits per slot)
    min_plain_bits: int = 30   # smallest scale a plaintext factor is encoded at
    waist_bootstrap: bool = True  # place bootstraps at single-ciphertext program points, sized to the next segment
    # per-op latency table (microseconds, index = level - 1) in the schema of profiled_B200_GPU.json / EarthDialect.cpp:130-180
    # ("latencyTableExact" or "latencyTable"); when given, the level a placed bootstrap returns to is the one that
    # minimises the estimated latency of the segment it feeds (DaCapo's objective), not merely the lowest feasible one
    cost_table: dict = None
    split_rotations: bool = True  # emit composite rotations as their NAF single-key steps (shared prefixes are CSE'd)
    fold_tolerance: float = 0.0


class _NeedBootstrap(Exception):
    pass


class _Ct:
    __slots__ = ("ssa", "level", "scale")

    def __init__(self, ssa, level, scale):
        self.ssa, self.level, self.scale = ssa, level, scale

    @property
    def bits(self):
        return math.log2(self.scale)


class Compiler:
    def __init__(self, graph: Graph, opt: Options = None):
        self.g = graph
        self.o = opt or Options()
        self.slots = 1 << (self.o.logN - 1)
        self.top = self.o.num_primes - 1
        self.q = prime_chain(self.o.logN, 60, self.o.num_primes)
        self.pool, self.pool_key = [], {}   # plaintext constant pool (.cst)
        self.ops = []                       # lowered SSA ops: (name, dst, a, b/imm)
        self.nssa = 0
        self.encodes = {}                   # (const id | -1, level, scale_bits) -> plaintext id
        self.stats = {}
        self.cse = {}
        self.min_level_seen = self.top
        self.greedy_forbidden = False
        self.est_us = 0.0                   # running latency estimate of the emitted ops (cost_table)

    # ---- latency model ---------------------------------------------------------------------------
    _COST_KEY = {"rotate": "earth.rotate_single", "mulcc": "earth.mul_double", "mulcp": "earth.mul_single",
                 "addcc": "earth.add_double", "addcp": "earth.add_single", "rescale": "earth.rescale_single",
                 "modswitch": "earth.modswitch_single", "negate": "earth.negate_single", "bootstrap": "earth.bootstrap_single"}

    def op_cost(self, name, imm, level):
        """Estimated microseconds of one lowered op whose RESULT is at `level`."""
        t = self.o.cost_table
        if not t or name not in self._COST_KEY:
            return 0.0
        row = t.get(self._COST_KEY[name]) or [0.0]
        src = level + 1 if name == "rescale" else level + imm if name == "modswitch" else imm if name == "bootstrap" else level
        c = row[min(max(src, 1), len(row)) - 1]
        if name == "rotate":  # SEAL rotates by NAF terms unless the step has a key of its own (a power of two)
            k = imm % self.slots
            k = min(k, self.slots - k)
            c *= max(1, bin(k).count("1") if k & (k - 1) else 1) if k else 0.05
        return c

    def naf_terms(self, step):
        """Non-adjacent form of a rotation step (SEAL util::naf order: increasing powers of two); a power of two is a
        single term.  Terms of +-slots are whole turns and are dropped like the runtime does."""
        sign, v, out, i = (-1 if step < 0 else 1), abs(step), [], 0
        while v:
            z = 2 - (v & 3) if v & 1 else 0
            v = (v - z) >> 1
            if z and (1 << i) != self.slots:
                out.append(sign * z * (1 << i))
            i += 1
        return out

    # ---- constants -------------------------------------------------------------------------------
    def tile(self, a):
        return np.resize(a, self.slots)   # the runtime tiles src[i % len] (SEAL_HEVM.cpp:259-261)

    def const_id(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        if arr.size > self.slots:
            arr = arr[:self.slots]
        if arr.size > 1 and np.all(arr == arr[0]):
            arr = arr[:1]
        key = (arr.size, arr.tobytes())
        if key not in self.pool_key:
            self.pool_key[key] = len(self.pool)
            self.pool.append(arr)
        return self.pool_key[key]

    def fold(self, fn, *ids):
        arrs = [self.pool[i] for i in ids]
        if all(a.size == 1 for a in arrs):
            return self.const_id(fn(*arrs))
        return self.const_id(fn(*[self.tile(a) for a in arrs]))

    # ---- emission helpers --------------------------------------------------------------------------
    def emit(self, name, srcs, imm=0, level=None, scale=None):
        # common-subexpression elimination: a value consumed several times is rescaled / bootstrapped /
        # level-dropped once, and identical traced ops are shared
        key = (name, tuple(srcs), imm)
        if name != "input":
            if key in self.cse:
                return self.cse[key]
        v = self._emit(name, srcs, imm, level, scale)
        if name != "input":
            self.cse[key] = v
            if level is not None and level < self.min_level_seen:
                self.min_level_seen = level
        return v

    def _emit(self, name, srcs, imm, level, scale):
        dst = self.nssa
        self.nssa += 1
        self.ops.append((name, dst, tuple(srcs), imm))
        self.stats[name] = self.stats.get(name, 0) + 1
        if level is not None:
            self.est_us += self.op_cost(name, imm, level)
        return _Ct(dst, level, scale)

    def valid(self, level, bits):
        return bits + self.o.margin_bits <= 60 * level - 1

    def plain(self, cid, level, scale_bits):
        key = (cid, level, int(scale_bits))
        if key not in self.encodes:
            self.encodes[key] = len(self.encodes)
        return self.encodes[key]

    def bootstrap(self, v, tgt=None):
        if self.greedy_forbidden:
            raise _NeedBootstrap()
        while v.bits > 105 and v.level >= 2:  # re-encoding must stay inside the encoder's 128-bit path
            v = self.emit("rescale", [v.ssa], 0, v.level - 1, v.scale / float(self.q[v.level - 1]))
        if not self.valid(v.level, v.bits):
            raise RuntimeError(f"cannot bootstrap: scale 2^{v.bits:.1f} does not fit level {v.level}")
        tgt = tgt or self.o.boot_target or self.top
        sb = int(v.bits + 1e-9)  # (int64) log2(scale) truncation of the wrapper (SEAL_HEVM.cpp:332)
        return self.emit("bootstrap", [v.ssa], tgt, tgt, 2.0 ** sb)

    def rescale(self, v):
        if v.level < 2:
            v = self.bootstrap(v)
        return self.emit("rescale", [v.ssa], 0, v.level - 1, v.scale / float(self.q[v.level - 1]))

    def normalize(self, v):
        W, R = self.o.waterline, self.o.rescale_bits
        while v.bits >= W + R - 0.5:
            v = self.rescale(v)
        return v

    def modswitch_to(self, v, level):
        if v.level == level:
            return v
        return self.emit("modswitch", [v.ssa], v.level - level, level, v.scale)

    def upscale(self, v, k):
        if not self.valid(v.level, v.bits + k):
            v = self.bootstrap(v)
        pt = self.plain(-1, v.level, k)
        return self.emit("mulcp", [v.ssa], pt, v.level, v.scale * 2.0 ** k)

    def mulcp(self, a, cid):
        W, R = self.o.waterline, self.o.rescale_bits
        a = self.normalize(a)
        # plaintext scale: R/2 = 30 bits while the product stays below the rescaling threshold (two plaintext
        # multiplications then cost exactly one level), otherwise just enough to land on waterline + R
        P = self.o.plain_bits
        k = P if a.bits + P <= W + R + 0.5 else max(self.o.min_plain_bits, int(round(W + R - a.bits)))
        if not self.valid(a.level, a.bits + k):
            a = self.normalize(self.bootstrap(a))
        pt = self.plain(cid, a.level, k)
        return self.emit("mulcp", [a.ssa], pt, a.level, a.scale * 2.0 ** k)

    def mulcc(self, a, b):
        same = a.ssa == b.ssa
        a = self.normalize(a)
        b = a if same else self.normalize(b)
        lv = min(a.level, b.level)
        for _ in range(3):  # refresh only the operand(s) sitting on the limiting (lowest) level
            if self.valid(lv, a.bits + b.bits):
                break
            if same:
                a = b = self.normalize(self.bootstrap(a))
            elif a.level <= b.level:
                a = self.normalize(self.bootstrap(a))
            else:
                b = self.normalize(self.bootstrap(b))
            lv = min(a.level, b.level)
        a, b = self.modswitch_to(a, lv), (None if same else self.modswitch_to(b, lv))
        if same:
            b = a
        return self.emit("mulcc", [a.ssa, b.ssa], 0, lv, a.scale * b.scale)

    def addcc(self, a, b):
        W, R = self.o.waterline, self.o.rescale_bits
        for _ in range(8):
            d = a.bits - b.bits
            if abs(d) <= 0.5:
                break
            hi, lo = (a, b) if d > 0 else (b, a)
            if abs(d) >= R - 0.5 and hi.bits >= W + R - 0.5:
                hi = self.rescale(hi)                      # about one prime (or more) apart: rescale the larger
            else:
                lo = self.upscale(lo, int(round(abs(d))))  # otherwise lift the smaller: Encode(ones, 2^k) + MulCP
            a, b = (hi, lo) if d > 0 else (lo, hi)
        lv = min(a.level, b.level)
        a, b = self.modswitch_to(a, lv), self.modswitch_to(b, lv)
        return self.emit("addcc", [a.ssa, b.ssa], 0, lv, b.scale)  # the wrapper takes the rhs scale (SEAL_HEVM.cpp:301)

    def addcp(self, a, cid):
        if a.bits > 105:   # keep the encoded plaintext coefficients inside the encoder's 128-bit path
            a = self.normalize(a)
        k = int(round(a.bits))
        pt = self.plain(cid, a.level, k)
        return self.emit("addcp", [a.ssa], pt, a.level, 2.0 ** k)  # lhs scale := plaintext scale (SEAL_HEVM.cpp:308)

    # ---- the pass ---------------------------------------------------------------------------------------
    def _operands(self, n):
        if n[0] in ("add", "sub", "mul"):
            return n[1:3]
        if n[0] in ("neg", "rot", "boot"):
            return n[1:2]
        return ()

    def _process(self, i):
        g, vals = self.g, self.vals
        n = g.nodes[i]
        k = n[0]
        if k == "input":
            v = self.emit("input", [], n[1], self.top, 2.0 ** self.o.waterline)
            self.inputs.append(v)
            vals[i] = v
        elif k == "const":
            vals[i] = self.const_id(g.consts[n[1]])
        elif k == "boot":
            vals[i] = vals[n[1]]  # placement is the compiler's job (the dacapo pipeline also drops manual ones)
        elif k == "neg":
            a = vals[n[1]]
            vals[i] = self.fold(np.negative, a) if isinstance(a, int) else self.emit("negate", [a.ssa], 0, a.level, a.scale)
        elif k == "rot":
            a, off = vals[n[1]], n[2]
            if isinstance(a, int):
                vals[i] = self.const_id(np.roll(self.tile(self.pool[a]), -off))
            elif off % self.slots == 0:
                vals[i] = a
            else:
                off = ((off + self.slots // 2) % self.slots) - self.slots // 2  # into [-slots/2, slots/2)
                # A step without a Galois key of its own is executed by the runtime as SEAL's rotate_internal does: one
                # key switch per NAF term, least significant first (Evaluator::rotate_internal, SEAL_HEVM.cpp:273).
                # Emitting those single-key steps explicitly yields the same ciphertext bit for bit and lets the CSE
                # share common prefixes between the many rotations of one value (convolution taps).
                v = a
                for term in (self.naf_terms(off) if self.o.split_rotations else [off]):
                    v = self.emit("rotate", [v.ssa], term, v.level, v.scale)
                vals[i] = v
        else:
            a, b = vals[n[1]], vals[n[2]]
            pa, pb = isinstance(a, int), isinstance(b, int)
            if pa and pb:
                fn = {"add": np.add, "sub": np.subtract, "mul": np.multiply}[k]
                vals[i] = self.fold(fn, a, b)
            elif k == "mul":
                if pa or pb:
                    c, p = (b, a) if pa else (a, b)
                    arr = self.pool[p]
                    vals[i] = c if np.all(arr == 1.0) else self.mulcp(c, p)
                else:
                    vals[i] = self.mulcc(a, b)
            else:
                if k == "sub":  # a - b = a + (-b)
                    b = self.fold(np.negative, b) if pb else self.emit("negate", [b.ssa], 0, b.level, b.scale)
                if pa or pb:
                    c, p = (b, a) if pa else (a, b)
                    vals[i] = c if not np.any(self.pool[p]) else self.addcp(c, p)
                else:
                    vals[i] = self.addcc(a, b)

    def _snapshot(self):
        return (len(self.ops), self.nssa, dict(self.stats), dict(self.cse), dict(self.encodes), dict(self.vals),
                len(self.pool), dict(self.pool_key), self.min_level_seen, self.est_us)

    def _restore(self, snap):
        nops, self.nssa, stats, cse, encodes, vals, npool, pool_key, self.min_level_seen, self.est_us = snap
        # copies: a snapshot may be restored several times and must stay pristine
        self.stats, self.cse, self.encodes, self.vals, self.pool_key = dict(stats), dict(cse), dict(encodes), dict(vals), dict(pool_key)
        del self.ops[nops:]
        del self.pool[npool:]

    def lower(self):
        """Forward pass over the traced nodes.  Bootstraps are placed at *waists* -- program points crossed by a
        single live ciphertext (layer boundaries) -- whenever a dry run of the next segment shows that the levels
        would otherwise run out inside it, and only up to the level that segment needs (cheaper ops everywhere);
        anything still short of levels inside a segment is bootstrapped greedily, as late as possible."""
        g = self.g
        self.vals, self.inputs, self.min_level_seen = {}, [], self.top
        reach, stack = set(), list(g.outputs)
        while stack:  # dead code elimination
            i = stack.pop()
            if i in reach:
                continue
            reach.add(i)
            stack.extend(self._operands(g.nodes[i]))
        order = [i for i, n in enumerate(g.nodes) if i in reach or n[0] == "input"]
        # cipher-ness and last use, to find the waists
        is_ct, last_use = {}, {}
        for i in order:
            n = g.nodes[i]
            is_ct[i] = n[0] == "input" or any(is_ct.get(s, False) for s in self._operands(n))
            for s in self._operands(n):
                last_use[s] = i
        for i in g.outputs:
            last_use[i] = len(g.nodes)
        waists, live, pos = [], 0, {i: k for k, i in enumerate(order)}
        ends = {}
        for i in order:
            if is_ct[i] and g.nodes[i][0] != "boot":
                ends.setdefault(last_use.get(i, i), []).append(i)
        live_set = set()
        for i in order:
            for v in ends.get(i, ()):   # values whose last use is node i die here
                live_set.discard(v)
            if is_ct[i] and g.nodes[i][0] != "boot" and last_use.get(i, i) > i:
                live_set.add(i)
            if len(live_set) == 1 and i in live_set and is_ct[i]:
                waists.append(i)
        self.waists = waists
        wset = set(waists) if self.o.waist_bootstrap else set()

        def is_waist(i):
            return i in wset and not isinstance(self.vals[i], int) and g.nodes[i][0] != "input"

        # ---- pass 1: greedy, but when the levels run out rewind to the most recent waist and bootstrap there ----
        base_snap = self._snapshot()
        placed = {}
        k, rewind = 0, None
        while k < len(order):
            i = order[k]
            self.greedy_forbidden = rewind is not None
            try:
                self._process(i)
                if k == len(order) - 1:
                    for o_ in g.outputs:
                        self.normalize(self.vals[o_])
            except _NeedBootstrap:
                snap, w, kw = rewind
                self._restore(snap)
                placed[w] = self.top
                self.greedy_forbidden = False
                self.vals[w] = self.bootstrap(self.vals[w], self.top)
                k, rewind = kw + 1, None
                continue
            if is_waist(i) and i not in placed:
                rewind = (self._snapshot(), i, k)
            k += 1
        self.greedy_forbidden = False
        # ---- pass 2: replay with the placement fixed; size every placed bootstrap to what its segment needs ----
        self._restore(base_snap)
        self.inputs = []
        marks = sorted(placed, key=lambda w: pos[w])
        nxt_mark = {w: (marks[t + 1] if t + 1 < len(marks) else None) for t, w in enumerate(marks)}
        k = 0
        while k < len(order):
            i = order[k]
            self._process(i)
            if i in placed:
                stop = nxt_mark[i]
                seg = order[k + 1:(pos[stop] + 1) if stop is not None else len(order)]
                v = self.vals[i]
                snap = self._snapshot()

                def boots_with(target, latency=False):
                    self._restore(snap)
                    t0 = self.est_us
                    self.vals[i] = self.bootstrap(v, target)
                    b0 = self.stats.get("bootstrap", 0)
                    for j in seg:
                        self._process(j)
                    if stop is None:
                        for o_ in g.outputs:
                            self.normalize(self.vals[o_])
                    else:
                        self.bootstrap(self.vals[stop], self.top)  # the next placed bootstrap must still be possible
                    return self.est_us - t0 if latency else self.stats.get("bootstrap", 0) - b0

                if self.o.cost_table:
                    # latency-driven: try every level, keep the cheapest feasible one (extra in-segment bootstraps are
                    # allowed when the cheaper low-level ops pay for them)
                    best_t, best_c = self.top, None
                    saved_bt = self.o.boot_target
                    for tgt in range(self.top, 1, -1):
                        try:
                            self.o.boot_target = tgt  # in-segment bootstraps return to the same level
                            c_ = boots_with(tgt, latency=True)
                        except (RuntimeError, _NeedBootstrap):
                            continue
                        finally:
                            self.o.boot_target = saved_bt
                        if best_c is None or c_ < best_c - 1e-9:
                            best_t, best_c = tgt, c_
                    self._restore(snap)
                    self.seg_targets = getattr(self, "seg_targets", {})
                    self.seg_targets[i] = best_t
                    self.o.boot_target = best_t
                    self.vals[i] = self.bootstrap(v, best_t)
                    k += 1
                    continue
                try:
                    base = boots_with(self.top)
                    lo, hi = 2, self.top
                    while lo < hi:
                        mid = (lo + hi) // 2
                        ok = False
                        try:
                            ok = boots_with(mid) <= base
                        except RuntimeError:
                            ok = False
                        if ok:
                            hi = mid
                        else:
                            lo = mid + 1
                except RuntimeError:
                    hi = self.top
                self._restore(snap)
                self.vals[i] = self.bootstrap(v, hi)
            k += 1
        outs = []
        for i in g.outputs:
            v = self.vals[i]
            if isinstance(v, int):
                raise ValueError("plaintext outputs are not supported")
            outs.append(self.normalize(v))
        return self.inputs, outs

    # ---- register allocation + emission --------------------------------------------------------------------
    def emit_program(self):
        inputs, outs = self.lower()
        last_use = {}
        for idx, (name, dst, srcs, imm) in enumerate(self.ops):
            for s in srcs:
                last_use[s] = idx
        for v in outs:
            last_use[v.ssa] = len(self.ops)  # live to the end
        p = asm.Program(init_level=self.top)
        reg_of, free = {}, []
        for v in inputs:
            reg_of[v.ssa] = p.arg(self.o.waterline, self.top)
        # plaintext registers: one Encode per (constant, level, scale)
        used_pool = {}
        for (cid, level, bits), pt in sorted(self.encodes.items(), key=lambda kv: kv[1]):
            reg = p.new_pt()
            assert reg == pt
            if cid >= 0 and cid not in used_pool:
                used_pool[cid] = p.const(self.pool[cid])
            p.encode(reg, -1 if cid < 0 else used_pool[cid], level, bits)
        opcode = {"rotate": asm.ROTATE, "negate": asm.NEGATE, "rescale": asm.RESCALE, "modswitch": asm.MODSWITCH,
                  "addcc": asm.ADDCC, "addcp": asm.ADDCP, "mulcc": asm.MULCC, "mulcp": asm.MULCP, "bootstrap": asm.BOOTSTRAP}
        for idx, (name, dst, srcs, imm) in enumerate(self.ops):
            if name == "input":
                continue
            src_regs = [reg_of[s] for s in srcs]
            for s in set(srcs):  # a source dying here frees its register first: dst may reuse it (runtime allows aliasing)
                if last_use.get(s) == idx and s not in [v.ssa for v in inputs]:
                    free.append(reg_of[s])
            if dst not in last_use:   # dead value (cannot happen after DCE, but stay safe)
                last_use[dst] = idx
            r = free.pop() if free else p.new_ct()
            reg_of[dst] = r
            if name in ("addcc", "mulcc"):
                p.emit(opcode[name], r, src_regs[0], src_regs[1])
            elif name == "rotate":
                p.rotate(r, src_regs[0], imm)
            else:
                p.emit(opcode[name], r, src_regs[0], imm)
        for v in outs:
            p.result(reg_of[v.ssa], int(round(v.bits)), v.level)
        return p


def compile_graph(graph: Graph, options: Options = None):
    c = Compiler(graph, options)
    prog = c.emit_program()
    return prog, c
