"""GPU tier: the real encrypted ResNet-20 (BASELINE.json configs[0]) from the committed fixture.

Program + constants were traced from the reference's examples/benchmarks/ResNet.py and compiled by
dacapo_b200.compiler (tests/golden/make_resnet_fixture.py); the expected logits come from the plaintext
torch model with the reference's weights.  Acceptance = the reference's own check (examples/tests/ResNet.py:
113-118): rms of (decrypted*32 - plaintext logits); the README reports 9.5e-4 for SEAL (README.md:187)."""
import ctypes as C
import time

import numpy as np
import pytest

from dacapo_b200 import fixtures
from util import make_vm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("waterline", [40, 35, 50])
def test_encrypted_resnet20_matches_plaintext_model(b200_lib, tmp_path, waterline):
    cst, hv, x, expected, meta = fixtures.resnet20_files(tmp_path)
    if waterline != 40:  # waterline sweep programs share the constant pool (tests/golden/make_resnet_fixture.py)
        hv = str(fixtures.FIXTURE / f"resnet20_w{waterline}.hevm")
    lib = b200_lib
    vm, _ = make_vm(lib, 15, 14)
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)
    f64p = C.POINTER(C.c_double)
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    lib.run(vm)                       # first run builds the schedule / CUDA graph
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    t = time.perf_counter()
    lib.run(vm)
    latency = time.perf_counter() - t
    out = np.zeros(1 << 14)
    lib.decrypt_result(vm, 0, out.ctypes.data_as(f64p))
    res = out[:meta["n_out"]] * meta["post_scale"]
    err = res - expected
    rms = float(np.sqrt(np.sum(err * err) / res.shape[-1]))
    print(f"encrypted ResNet-20 (waterline {waterline}): run() {latency:.3f}s rms {rms:.3e} argmax {int(np.argmax(res))} vs {int(np.argmax(expected))}")
    assert np.argmax(res) == np.argmax(expected)
    assert rms < 5e-3, rms


def test_encrypted_resnet20_at_2_16_slots(b200_lib, tmp_path):
    """BASELINE.json configs[2]: the benchmark's own default packing, nt = 2^16 slots => N = 2^17 (14 x 60-bit primes),
    keys and 1 882 plaintexts resident in HBM, bootstraps (decrypt + re-encrypt) on the device."""
    cst, hv, x, expected, meta = fixtures.resnet20_files(tmp_path, "_nt16")
    lib = b200_lib
    vm, _ = make_vm(lib, meta["logN"], 14)
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)
    f64p = C.POINTER(C.c_double)
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    lib.run(vm)
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    t = time.perf_counter()
    lib.run(vm)
    latency = time.perf_counter() - t
    out = np.zeros(meta["slots"])
    lib.decrypt_result(vm, 0, out.ctypes.data_as(f64p))
    res = out[:meta["n_out"]] * meta["post_scale"]
    rms = float(np.sqrt(np.mean((res - expected) ** 2)))
    print(f"encrypted ResNet-20, nt = 2^16 (N = 2^17): run() {latency:.3f}s rms {rms:.3e} argmax {int(np.argmax(res))} vs {int(np.argmax(expected))}")
    assert np.argmax(res) == np.argmax(expected)
    assert rms < 5e-3, rms
