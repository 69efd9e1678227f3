"""CPU tier: the restated reference compiler passes (dacapo_b200.earth = PARS pipeline, dacapo_b200.dacapo = DaCapo
bootstrap planner; reference: lib/Dialect/Earth/Transforms/*, tools/optimizer.cpp:380-480) produce HEVM programs that
decrypt to what plain numpy computes on the CPU oracle, on a small ring."""
import ctypes as C

import numpy as np
import pytest

from dacapo_b200 import dacapo, earth, frontend as hc
from util import make_vm

LOGN, NPR = 11, 8
f64p = C.POINTER(C.c_double)
N2 = 1 << (LOGN - 1)
W_ = np.linspace(-1, 1, N2)


def params():
    lat = {k: [0] + [10 * (l + 1) for l in range(8)] for k in ("earth.add_single", "earth.add_double", "earth.mul_single", "earth.negate_single",
                                                                  "earth.modswitch_single", "earth.upscale_single")}
    lat["earth.rotate_single"] = [0] + [100 * (l + 1) for l in range(8)]
    lat["earth.mul_double"] = [0] + [120 * (l + 1) for l in range(8)]
    lat["earth.rescale_single"] = [0] + [30 * (l + 1) for l in range(8)]
    lat["earth.bootstrap_single"] = [0] + [500] * 8
    return earth.Params(waterline=40, output_val=10, level_upper=7, level_lower=2, boot_upper=7, boot_lower=2, poly_degree=1 << LOGN, latency=lat)


def run_on(lib, prog, inputs, tmp_path):
    vm, _ = make_vm(lib, LOGN, NPR)
    cst, hv = tmp_path / "c.cst", tmp_path / "c.hevm"
    prog.save(cst, hv)
    lib.load(vm, str(cst).encode(), str(hv).encode())
    lib.preprocess(vm)
    for i, x in enumerate(inputs):
        lib.encrypt(vm, i, np.ascontiguousarray(x).ctypes.data_as(f64p), N2)
    lib.run(vm)
    out = np.zeros((lib.getResLen(vm), N2))
    for i in range(out.shape[0]):
        lib.decrypt_result(vm, i, out[i].ctypes.data_as(f64p))
    return out


def build(depth, manual):
    hc.reset()

    @hc.func("c")
    def f(x):
        y = x * W_ + 0.25
        for d in range(depth):
            y = y * y * 0.4 - x.rotate(d + 1) * 0.1 + y.rotate(3) * 0.01
            if manual and d % 2 == 1:
                y = hc.bootstrap(y)
        return y

    return hc.save()


def ref(x, depth):
    y = x * W_ + 0.25
    for d in range(depth):
        y = y * y * 0.4 - np.roll(x, -(d + 1)) * 0.1 + np.roll(y, -3) * 0.01
    return y


def test_naf_and_rotation_cost():
    assert earth.naf(12285) == [1, -4, 4096, 8192][:0] or sum(earth.naf(12285)) == 12285
    assert sum(earth.naf(-15360)) == -15360 and all(abs(t) & (abs(t) - 1) == 0 for t in earth.naf(-15360))
    P = params()
    r = earth.V("rot", [earth.V("arg")], 3)
    r.level = 2
    assert earth.op_latency(r, P, 7) == 2 * P.latency["earth.rotate_single"][5]  # NAF(3) = {-1, 4}: two key switches at CKKS level 5


def test_pars_rules_on_a_product_chain():
    """ProactiveRescaling: ct x pt puts the plaintext at the waterline; a ct x ct whose scales sum above 2W + Rf first
    brings its operands down to the waterline (upscale + rescale); a product is rescaled while a whole factor is spare."""
    hc.reset()

    @hc.func("c,c")
    def f(x, y):
        t = x * y            # 80
        u = t * t            # operands 80 + 80 > 140: both go to 40 first
        return u * 0.5 + x

    g = hc.save()
    P = params()
    fn = earth.from_graph(g)
    earth.pars(fn, P)
    muls = [o for o in fn.ops if o.kind == "mul"]
    assert [m.scale for m in muls] == [80, 80, 80 + 40]
    assert muls[1].ins[0].kind == "rescale" and muls[1].ins[0].ins[0].kind == "upscale" and muls[1].ins[0].scale == 40
    assert muls[2].ins[1].kind == "const" and (muls[2].ins[1].scale, muls[2].ins[1].level) == (40, muls[2].ins[0].level)
    earth.early_modswitch(fn, P)
    earth.canonicalize(fn, P)
    earth.verify(fn, P)
    # refineReturnValues: every result sits at the lowest level that still holds scale + output_val bits
    Rf = P.rescaling_factor
    assert all(fn.init_level * Rf - (r.level * Rf + r.scale + P.output_val) < Rf for r in fn.rets)


def test_pars_pipeline_on_oracle(oracle_lib, tmp_path):
    x = np.random.default_rng(0).uniform(-1, 1, N2)
    prog, fn = earth.compile_pars(build(2, False), params())
    assert "boot" not in earth.op_counts(fn)
    assert np.max(np.abs(run_on(oracle_lib, prog, [x], tmp_path)[0] - ref(x, 2))) < 1e-6
    prog, fn = earth.compile_pars(build(8, True), params())
    assert earth.op_counts(fn)["boot"] == 4
    assert np.max(np.abs(run_on(oracle_lib, prog, [x], tmp_path)[0] - ref(x, 8))) < 1e-6
    with pytest.raises(earth.PassFailed):  # eight squarings do not fit seven levels without a bootstrap
        earth.compile_pars(build(8, False), params())


def test_dacapo_planner_places_bootstraps(oracle_lib, tmp_path):
    """The manual bootstraps are removed (RemoveBootstrap) and the planner places its own; the program must still be right."""
    x = np.random.default_rng(1).uniform(-1, 1, N2)
    P = params()
    prog, fn, rep = dacapo.compile_dacapo(build(8, True), P)
    assert rep["bootstraps"] >= 3 and rep["selected_set"] >= 1 and rep["candidates"] > 2
    assert abs(rep["estimated_latency_final_s"] - earth.latency(fn, P) / 1e6) < 1e-12
    assert np.max(np.abs(run_on(oracle_lib, prog, [x], tmp_path)[0] - ref(x, 8))) < 1e-6
    # an expensive bootstrap makes the planner use fewer of them than a cheap one
    cheap, dear = params(), params()
    cheap.latency["earth.bootstrap_single"] = [0] + [1] * 8
    dear.latency["earth.bootstrap_single"] = [0] + [10 ** 7] * 8
    n_cheap = dacapo.compile_dacapo(build(8, False), cheap)[2]["bootstraps"]
    n_dear = dacapo.compile_dacapo(build(8, False), dear)[2]["bootstraps"]
    assert n_dear <= n_cheap


def test_smu_partition():
    """ScaleManagementUnit: the three addends of a sum of products share a unit with the sum; a multiply starts a new one."""
    hc.reset()

    @hc.func("c")
    def f(x):
        s = x.rotate(1) + x.rotate(2) + x.rotate(3)
        return s * s

    fn = earth.from_graph(hc.save())
    smu = dacapo.smu_ids(fn)
    adds = [o for o in fn.ops if o.kind == "add"]
    mul = [o for o in fn.ops if o.kind == "mul"][0]
    assert smu[id(adds[0])] == smu[id(adds[1])]
    assert smu[id(mul)] != smu[id(adds[1])]
