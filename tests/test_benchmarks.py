"""The reference's small benchmarks (examples/benchmarks: SobelFilter, HarrisCornerDetection, LinearRegression,
PolynomialRegression, Multivariate, MLP) end to end: traced from the unmodified reference scripts with
dacapo_b200.frontend, compiled by dacapo_b200.compiler (tests/golden/make_bench_fixtures.py), executed through the
reference's C ABI (load / preprocess / encrypt / run / decrypt_result) and compared with the plaintext evaluation of
the traced graph -- the check of examples/tests/<name>.py (rms of decrypted - expected)."""
import ctypes as C
import json
import lzma
from pathlib import Path

import numpy as np
import pytest

from util import make_vm

BENCH = Path(__file__).resolve().parent / "golden" / "bench"
NAMES = ["SobelFilter", "HarrisCornerDetection", "LinearRegression", "PolynomialRegression", "Multivariate", "MLP"]
f64p = C.POINTER(C.c_double)


def run_benchmark(lib, name, tmp_path):
    d = BENCH / name
    meta = json.loads((d / "meta.json").read_text())
    io = np.load(d / "io.npz")
    cst, hv = tmp_path / "p.cst", tmp_path / "p.hevm"
    cst.write_bytes(lzma.decompress((d / "prog.cst.xz").read_bytes()))
    hv.write_bytes((d / "prog.hevm").read_bytes())
    vm, _ = make_vm(lib, 15, 14)
    lib.load(vm, str(cst).encode(), str(hv).encode())
    lib.preprocess(vm)
    assert lib.getArgLen(vm) == meta["n_inputs"] and lib.getResLen(vm) == meta["n_outputs"]
    for i, x in enumerate(io["inputs"]):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib.encrypt(vm, i, x.ctypes.data_as(f64p), x.size)
    lib.run(vm)
    res = np.zeros((meta["n_outputs"], 1 << 14))
    for i in range(meta["n_outputs"]):
        lib.decrypt_result(vm, i, res[i].ctypes.data_as(f64p))
    err = res - io["expected"]
    return float(np.sqrt(np.mean(err * err))), float(np.max(np.abs(err))), meta


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_reference_benchmark_on_gpu(b200_lib, tmp_path, name):
    rms, worst, meta = run_benchmark(b200_lib, name, tmp_path)
    print(f"{name}: rms {rms:.2e} max {worst:.2e} (|expected| <= {meta['max_abs_expected']:.3g})")
    assert rms < 1e-4 * max(1.0, meta["max_abs_expected"]), (name, rms)


def test_sobel_filter_on_the_cpu_oracle(oracle_lib, tmp_path):
    """Same chain on the CPU restatement (one small benchmark: full-size keys take ~15 s to generate on the host)."""
    rms, worst, meta = run_benchmark(oracle_lib, "SobelFilter", tmp_path)
    assert rms < 1e-4 * max(1.0, meta["max_abs_expected"]), rms
