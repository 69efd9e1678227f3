"""CPU tier: tracing front end + compiler (dacapo_b200.frontend / compiler) -> HEVM program -> oracle.

The compiled program must decrypt to what plain numpy computes; this exercises scale management
(lazy rescale, upscale, modswitch alignment), automatic bootstrap insertion, constant folding and
register reuse, on a small ring so that the CPU oracle runs it in seconds."""
import ctypes as C

import numpy as np
import pytest

from dacapo_b200 import compiler, frontend as hc
from dacapo_b200 import hevm_asm as asm
from util import make_vm

LOGN, NPR = 11, 8  # 1024 slots, 7 data levels
f64p = C.POINTER(C.c_double)


def run_on(lib, prog, inputs, tmp_path):
    vm, _ = make_vm(lib, LOGN, NPR)
    cst, hv = tmp_path / "c.cst", tmp_path / "c.hevm"
    prog.save(cst, hv)
    lib.load(vm, str(cst).encode(), str(hv).encode())
    lib.preprocess(vm)
    n = 1 << (LOGN - 1)
    for i, x in enumerate(inputs):
        lib.encrypt(vm, i, np.ascontiguousarray(x).ctypes.data_as(f64p), n)
    lib.run(vm)
    out = np.zeros((lib.getResLen(vm), n))
    for i in range(out.shape[0]):
        lib.decrypt_result(vm, i, out[i].ctypes.data_as(f64p))
    return out


def test_polynomial_with_rotations(oracle_lib, tmp_path):
    hc.reset()
    n = 1 << (LOGN - 1)
    w = np.linspace(-1, 1, n)

    @hc.func("c")
    def f(x):
        y = x * w + 0.25            # ct*pt, ct+scalar
        z = y * y - x               # ct*ct, sub
        acc = hc.Empty()
        for k in (1, -3, 17):       # Empty accumulation + rotations
            acc = acc + z.rotate(k) * 0.5
        return acc * acc + z * 1.0  # mul by one is folded

    g = hc.save()
    prog, c = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR))
    assert c.stats.get("mulcc", 0) == 2 and c.stats.get("rotate", 0) == 3
    x = np.random.default_rng(0).uniform(-1, 1, n)
    out = run_on(oracle_lib, prog, [x], tmp_path)[0]
    y = x * w + 0.25
    z = y * y - x
    acc = sum(np.roll(z, -k) * 0.5 for k in (1, -3, 17))
    assert np.max(np.abs(out - (acc * acc + z))) < 1e-4


def test_deep_chain_inserts_bootstraps(oracle_lib, tmp_path):
    hc.reset()

    @hc.func("c")
    def f(x):
        y = x
        for i in range(12):         # 12 multiplicative levels > 7 available
            y = y * y * 0.9 + 0.05
        return y

    g = hc.save()
    prog, c = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR))
    assert c.stats.get("bootstrap", 0) >= 1
    n = 1 << (LOGN - 1)
    x = np.random.default_rng(1).uniform(-1, 1, n)
    out = run_on(oracle_lib, prog, [x], tmp_path)[0]
    y = x.copy()
    for i in range(12):
        y = y * y * 0.9 + 0.05
    assert np.max(np.abs(out - y)) < 1e-3


def test_constant_folding_and_residual_alignment(oracle_lib, tmp_path):
    hc.reset()
    n = 1 << (LOGN - 1)

    @hc.func("c")
    def f(x):
        k = hc.Plain([2.0]) * hc.Plain([0.25]) + hc.Plain([0.5])     # plain (op) plain -> 1.0
        a = x * k                                                       # folded away
        b = (x * [0.5, -0.5]) * (x * 0.3)                               # scale 2^80-ish after products
        return a + b + (x.rotate(2) - 0.125)                            # operands at different scales / levels

    g = hc.save()
    prog, c = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR))
    x = np.random.default_rng(2).uniform(-1, 1, n)
    out = run_on(oracle_lib, prog, [x], tmp_path)[0]
    ref = x + (x * np.resize([0.5, -0.5], n)) * (x * 0.3) + (np.roll(x, -2) - 0.125)
    assert np.max(np.abs(out - ref)) < 1e-4
    text = asm.disassemble(prog)
    assert "rotate" in text and "mulcc" in text


def test_latency_driven_bootstrap_levels_and_graph_evaluator(oracle_lib, tmp_path):
    """Options.cost_table (the measured profile's schema) picks the level every placed bootstrap returns to; the
    program still computes frontend.evaluate(graph).  With a table in which bootstraps are cheap and key switches
    steep in the level, the planner bootstraps more often and keeps rotations low."""
    hc.reset()
    n = 1 << (LOGN - 1)

    @hc.func("c")
    def f(x):
        y = x
        for i in range(6):
            s = y + y.rotate(1 << i) * 0.5          # rotation-heavy layer ...
            y = s * s * 0.45 + 0.05                 # ... followed by a ct x ct level
        return y

    g = hc.save()
    top = NPR - 1
    steep = {"earth.rotate_single": [10.0 * (l + 1) ** 2 for l in range(top)], "earth.mul_double": [10.0 * (l + 1) ** 2 for l in range(top)],
             "earth.mul_single": [1.0] * top, "earth.add_double": [1.0] * top, "earth.add_single": [1.0] * top,
             "earth.rescale_single": [2.0] * top, "earth.modswitch_single": [0.5] * top, "earth.negate_single": [1.0] * top,
             "earth.bootstrap_single": [5.0] * top}
    base, cb = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR))
    prog, c = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR, cost_table=steep))
    assert c.stats.get("bootstrap", 0) >= cb.stats.get("bootstrap", 0)
    x = np.random.default_rng(5).uniform(-1, 1, n)
    expected = hc.evaluate(g, [x], n)[0]
    for p in (base, prog):
        out = run_on(oracle_lib, p, [x], tmp_path)[0]
        assert np.max(np.abs(out - expected)) < 1e-3


def test_composite_rotation_is_split_into_naf_steps():
    """rotate(k) without a key of its own is emitted as SEAL's NAF steps, least significant first (Evaluator::rotate_internal)."""
    hc.reset()

    @hc.func("c")
    def f(x):
        return x.rotate(7) + x.rotate(-3)

    g = hc.save()
    prog, c = compiler.compile_graph(g, compiler.Options(logN=LOGN, num_primes=NPR))
    steps = [(r if r < 32768 else r - 65536) for oc, d, l, r in prog.ops if oc == asm.ROTATE]
    assert sorted(steps) == sorted([-1, 8, 1, -4])  # 7 = -1 + 8, -3 = +1 - 4
    cmp_ = compiler.Compiler(g, compiler.Options(logN=LOGN, num_primes=NPR))
    assert cmp_.naf_terms(7) == [-1, 8] and cmp_.naf_terms(-3) == [1, -4] and cmp_.naf_terms(64) == [64]


def test_profile_json_schema_with_throughput_table():
    """dacapo_b200/profile.py: the emitted cost profile keeps the reference loader's schema (EarthDialect.cpp:130-180: integer
    microseconds >= 1, entry k <=> level k + 1, rescale / modswitch rows padded at level 1) and carries the under-load table
    (`latencyTableThroughput`) with the same keys and indexing; the committed profiled_B200_GPU.json has both, and the
    throughput row of a key switch is below its lone-op latency at every level."""
    import json
    from dacapo_b200 import profile
    top = 13
    ops = ("rotate", "mulcc", "rescale", "modswitch", "addcc", "addcp", "mulcp", "negate", "bootstrap")
    lone = {op: {l: 10.0 * l + 0.4 for l in range(1 if op not in ("rescale", "modswitch") else 2, top + 1)} for op in ops}
    tp = {op: {l: v / 4 for l, v in lv.items()} for op, lv in lone.items()}
    prof = profile.profile_json(lone, 15, 14, tp)
    keys = {"earth.rotate_single", "earth.rescale_single", "earth.modswitch_single", "earth.add_single", "earth.add_double",
            "earth.mul_single", "earth.mul_double", "earth.negate_single", "earth.bootstrap_single"}
    for tab in ("latencyTable", "latencyTableExact", "latencyTableThroughput"):
        assert set(prof[tab]) == keys and all(len(v) == top for v in prof[tab].values()), tab
    assert all(isinstance(x, int) and x >= 1 for v in prof["latencyTable"].values() for x in v)
    assert prof["latencyTable"]["earth.rotate_single"][3] == 41 and prof["latencyTableThroughput"]["earth.rotate_single"][3] == 10.1
    assert prof["latencyTableThroughput"]["earth.rescale_single"][0] == prof["latencyTableThroughput"]["earth.rescale_single"][1]
    assert prof["polynomialDegree"] == 1 << 15 and prof["levelUpperBound"] == top and "noiseTable" in prof
    assert "latencyTableThroughput" not in profile.profile_json(lone, 15, 14)
    from pathlib import Path
    committed = json.loads((Path(__file__).resolve().parent.parent / "profiled_B200_GPU.json").read_text())
    for k in ("earth.rotate_single", "earth.mul_double", "earth.bootstrap_single"):
        assert len(committed["latencyTableThroughput"][k]) == top
        assert all(t < e for t, e in zip(committed["latencyTableThroughput"][k], committed["latencyTableExact"][k])), k
