"""Synthetic HEVM programs for measurement.

`resnet20_opmix()` builds the ResNet-20 *op-mix replay* of SURVEY.md section 8(d): the reference
compiler (hecate-opt, MLIR) cannot be built in this environment, so no compiled
`optimized/dacapo/ResNet.40._hecate_ResNet.hevm` exists; the fallback the survey defines is a program
with the op counts measured by tracing `examples/benchmarks/ResNet.py` at nt = 2^14:

    rotate 2 510 (3 910 key-switch steps after SEAL's NAF split), ct x ct 361, ct x pt 4 822,
    ct + ct 6 074, ct + pt 153, negate 133, bootstrap 19, 4 975 distinct plaintext constants

arranged like the network: 19 bootstrap-delimited segments; each segment walks the modulus chain
13 -> 2 through convolution-shaped blocks (independent rotations of one input, plaintext multiplies,
an accumulation tree, one rescale) and polynomial-activation-shaped blocks (ct x ct products with
upscale-multiplies and rescales).  Scale management ops (rescale / upscale-mulcp) are added the way
PARS does (rescale after every multiplication level, `UpscaleToMulcp.cpp:52-72` lowering).
It is a throughput / latency proxy, NOT the compiled ResNet: the numerical output is meaningless.
"""
import numpy as np

from . import hevm_asm as asm

TOP = 13
COUNTS = dict(rotate=2510, ks_steps=3910, mulcc=361, mulcp=4822, addcc=6074, addcp=153, negate=133, bootstrap=19,
              constants=4975)


def _naf_terms(v, half):
    res, sign, v, i = [], v < 0, abs(v), 0
    while v:
        z = (2 - (v & 3)) if v & 1 else 0
        v = (v - z) >> 1
        if z:
            res.append((-z if sign else z) * (1 << i))
        i += 1
    return [t for t in res if abs(t) != half]


def _offset_pool(rng, n_rot, n_steps, slots=1 << 14):
    """n_rot offsets whose NAF decompositions total ~n_steps key-switch steps (powers of two have keys)."""
    singles = [s * (1 << k) for k in range(0, 14) for s in (1, -1)]
    doubles = []
    for a in range(0, 13):
        for b in range(a + 2, 14):
            for sa in (1, -1):
                v = (1 << b) + sa * (1 << a)
                if abs(v) < slots // 2:
                    doubles.append(v)
                    doubles.append(-v)
    n_double = max(0, min(n_rot, n_steps - n_rot))
    offs = [doubles[i] for i in rng.integers(0, len(doubles), n_double)] + \
           [singles[i] for i in rng.integers(0, len(singles), n_rot - n_double)]
    rng.shuffle(offs)
    return [int(o) for o in offs]


def resnet20_opmix(seed=0):
    rng = np.random.default_rng(seed)
    p = asm.Program(init_level=TOP)
    x = p.arg(40, TOP)
    offsets = _offset_pool(rng, COUNTS["rotate"], COUNTS["ks_steps"])
    budget = {k: COUNTS[k] for k in ("rotate", "mulcc", "mulcp", "addcc", "addcp", "negate")}
    n_const = [0]
    stats = {k: 0 for k in ("rotate", "ks_steps", "mulcc", "mulcp", "addcc", "addcp", "negate", "rescale", "bootstrap", "encode")}
    per_level = {}

    def count(op, lvl, n=1):
        stats[op] = stats.get(op, 0) + n
        per_level[(op, lvl)] = per_level.get((op, lvl), 0) + n

    def new_const(level, scale_bits):
        pt = p.new_pt()
        if n_const[0] < COUNTS["constants"]:
            c = p.const(rng.uniform(-1, 1, 8) / 64.0)  # contractive taps keep every value in (-1, 1)
            n_const[0] += 1
        else:
            c = int(rng.integers(0, len(p.constants)))
        p.encode(pt, c, level, scale_bits)
        stats["encode"] += 1
        return pt

    # a small pool of ciphertext registers reused like ReuseBuffer does
    regs = [p.new_ct() for _ in range(40)]
    cur = regs[0]
    p.rotate(cur, x, 0)
    free = regs[1:]
    seg_n = COUNTS["bootstrap"]
    for seg in range(seg_n):
        left = seg_n - seg
        want = {k: -(-budget[k] // left) for k in budget}  # ceil share of what is left
        lvl = TOP
        # --- 6 convolution-shaped blocks, one level each -------------------------------------------
        n_blocks = 6
        for b in range(n_blocks):
            nrot = max(1, want["rotate"] // n_blocks)
            nmulp = max(nrot, want["mulcp"] // n_blocks - 2)
            acc = None
            nadd_extra = max(0, want["addcc"] // n_blocks - (nmulp - 1) - 3)
            for i in range(nmulp):
                r = free.pop()
                if i < nrot and budget["rotate"] > 0:
                    off = offsets.pop() if offsets else 1
                    p.rotate(r, cur, off)
                    budget["rotate"] -= 1
                    count("rotate", lvl)
                    count("ks_steps", lvl, max(1, len(_naf_terms(off, 1 << 13))))
                    src = r
                else:
                    src = cur
                p.emit(asm.MULCP, r, src, new_const(lvl, 60))
                budget["mulcp"] -= 1
                count("mulcp", lvl)
                if acc is None:
                    acc = r
                else:  # running accumulation (the q-loop of MultParConvBN), the tap register is recycled
                    p.emit(asm.ADDCC, acc, acc, r)
                    budget["addcc"] -= 1
                    count("addcc", lvl)
                    free.append(r)
            for _ in range(nadd_extra):  # bias / mask / residual additions of the traced graph
                p.emit(asm.ADDCC, acc, acc, cur)
                budget["addcc"] -= 1
                count("addcc", lvl)
            if budget["addcp"] > 0:
                p.emit(asm.ADDCP, acc, acc, new_const(lvl, 100))
                budget["addcp"] -= 1
                count("addcp", lvl)
            if budget["negate"] > 0:
                p.emit(asm.NEGATE, acc, acc)
                budget["negate"] -= 1
                count("negate", lvl)
            p.emit(asm.RESCALE, acc, acc)
            count("rescale", lvl)
            free.append(cur)
            cur = acc
            lvl -= 1
        # --- activation-shaped block: ct x ct products over the remaining levels -----------------------
        nmulcc = want["mulcc"]
        depth = lvl - 2  # stop at level 2 (bootstrapLevelLowerBound)
        per = -(-nmulcc // max(1, depth))
        for d in range(depth):
            outs = []
            for i in range(per):
                if budget["mulcc"] <= 0:
                    break
                r = free.pop()
                p.emit(asm.MULCC, r, cur, cur)
                budget["mulcc"] -= 1
                count("mulcc", lvl)
                ones = p.new_pt()
                p.encode(ones, -1, lvl, 20)  # upscale lowering: Encode(ones, 2^20) + MulCP
                stats["encode"] += 1
                p.emit(asm.MULCP, r, r, ones)
                budget["mulcp"] -= 1
                count("mulcp", lvl)
                outs.append(r)
            if not outs:
                break
            while len(outs) > 1:
                a, b2 = outs.pop(), outs.pop()
                p.emit(asm.ADDCC, a, a, b2)
                budget["addcc"] -= 1
                count("addcc", lvl)
                free.append(b2)
                outs.insert(0, a)
            p.emit(asm.RESCALE, outs[0], outs[0])
            count("rescale", lvl)
            free.append(cur)
            cur = outs[0]
            lvl -= 1
        # --- bootstrap back to the top level -----------------------------------------------------------
        r = free.pop()
        p.emit(asm.BOOTSTRAP, r, cur, TOP)
        count("bootstrap", lvl)
        free.append(cur)
        cur = r
    p.result(cur, 40, TOP)
    return p, stats, per_level
