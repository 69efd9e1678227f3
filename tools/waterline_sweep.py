"""Waterline sweep 30..50 of the encrypted ResNet-20 (BASELINE.json configs[3]) on the GPU: the committed programs
tests/golden/resnet20/resnet20[_wNN].hevm share one constant pool; prints latency of run() and the rms against the
plaintext model for each, as one JSON object.  usage: python tools/waterline_sweep.py [out.json]"""
import ctypes as C, json, os, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding
import fixtures
from util import make_vm

lib = _binding.bind(_binding.B200_LIB)
tmp = tempfile.mkdtemp()
cst, hv, x, expected, meta = fixtures.resnet20_files(tmp)
f64p = C.POINTER(C.c_double)
out = {"what": "encrypted ResNet-20, nt = 2^14, N = 2^15, 14 x 60-bit primes; programs compiled by dacapo_b200.compiler with "
               "bootstrap levels chosen against profiled_B200_GPU.json", "waterlines": {}}
vm, _ = make_vm(lib, 15, 14)
for W in (30, 35, 40, 45, 50):
    prog = hv if W == 40 else str(REPO / "tests" / "golden" / "resnet20" / f"resnet20_w{W}.hevm")
    if not os.path.isfile(prog):
        continue
    lib.load(vm, cst.encode(), prog.encode())
    t0 = time.perf_counter(); lib.preprocess(vm); pre = time.perf_counter() - t0
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    lib.run(vm)
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    t0 = time.perf_counter(); lib.run(vm); lat = time.perf_counter() - t0
    res = np.zeros(1 << 14)
    lib.decrypt_result(vm, 0, res.ctypes.data_as(f64p))
    r = res[:meta["n_out"]] * meta["post_scale"]
    ops = meta["lowered_ops"] if W == 40 else meta.get("waterline_sweep", {}).get(str(W), {}).get("lowered_ops", {})
    out["waterlines"][str(W)] = {"run_latency_s": round(lat, 4), "rms": float(np.sqrt(np.mean((r - expected) ** 2))),
                                 "argmax_ok": bool(np.argmax(r) == np.argmax(expected)), "preprocess_s": round(pre, 2),
                                 "bootstraps": ops.get("bootstrap"), "rescales": ops.get("rescale"), "mulcp": ops.get("mulcp")}
print(json.dumps(out))
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
