"""End-to-end example in the style of the reference's examples/tests/*.py (e.g. MLP.py:52-85,
ResNet.py:85-118): assemble a program, run it through the HEVM driver on the B200 backend, compare
with plain numpy and print latency / rms with HEVM.printer.

    python examples/poly_regression.py dacapo 40 B200 GPU

The reference produces its .hevm/.cst files with hecate-opt; that compiler is not part of this
repository, so the program (a degree-3 polynomial regression with a rotation-based moving average)
is assembled with dacapo_b200.hevm_asm.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dacapo_b200 as hc  # noqa: E402
from dacapo_b200 import hevm_asm as asm  # noqa: E402


def build(tmp):
    p = asm.Program(init_level=13)
    x = p.arg(50, 6)
    t, u, v = p.new_ct(), p.new_ct(), p.new_ct()
    w1, w3, up = p.new_pt(), p.new_pt(), p.new_pt()
    p.emit(asm.MULCC, t, x, x)               # x^2                 level 6, scale 2^100
    p.emit(asm.RESCALE, t, t)                #                     level 5, scale ~2^40
    p.emit(asm.MODSWITCH, t, t, 1)           #                     level 4
    p.emit(asm.MODSWITCH, u, x, 1)           # x                   level 5, scale 2^50
    p.encode(w3, p.const([-0.2]), 5, 50)
    p.emit(asm.MULCP, u, u, w3)              # -0.2 x              level 5, scale 2^100
    p.emit(asm.RESCALE, u, u)                #                     level 4, scale ~2^40
    p.emit(asm.MULCC, t, t, u)               # -0.2 x^3            level 4, scale ~2^80
    p.encode(w1, p.const([0.75]), 6, 50)
    p.emit(asm.MULCP, v, x, w1)              # 0.75 x              level 6, scale 2^100
    p.emit(asm.RESCALE, v, v)                #                     level 5, scale ~2^40
    p.emit(asm.MODSWITCH, v, v, 1)           #                     level 4
    p.encode(up, -1, 4, 40)                  # upscale lowering (UpscaleToMulcp.cpp:52-72): ones at 2^40
    p.emit(asm.MULCP, v, v, up)              #                     level 4, scale ~2^80
    p.emit(asm.ADDCC, t, t, v)               # y = 0.75 x - 0.2 x^3
    p.rotate(u, t, 1)
    p.emit(asm.ADDCC, t, t, u)               # y[i] + y[i+1]
    p.result(t, 80, 4)
    cst, hv = Path(tmp) / "_hecate_poly.cst", Path(tmp) / "optimized" / "dacapo"
    hv.mkdir(parents=True, exist_ok=True)
    hv = hv / "poly.40._hecate_poly.hevm"
    p.save(cst, hv)
    return str(cst), str(hv)


def reference(x):
    y = 0.75 * x - 0.2 * x ** 3
    return y + np.roll(y, -1)


if __name__ == "__main__":
    import tempfile
    hc.setLibnHW(sys.argv)
    hevm = hc.HEVM()
    cst, hv = build(tempfile.mkdtemp())
    hevm.load(cst, hv)
    x = np.random.default_rng(100).uniform(-1, 1, hevm.slots)
    hevm.setInput(0, x)
    timer = time.perf_counter_ns()
    hevm.run()
    timer = time.perf_counter_ns() - timer
    res = hevm.getOutput()[0]
    err = res - reference(x)
    rms = np.sqrt(np.sum(err * err) / res.shape[-1])
    hevm.printer(timer / 1e9, rms)
