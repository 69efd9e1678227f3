"""Kernel launches and latency of one run() of the ResNet-20 fixture (env switches: HEVM_CHAIN, HEVM_FUSE, HEVM_STREAMS)."""
import ctypes as C, os, sys, tempfile, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding
import fixtures
from util import make_vm
lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
cst, hv, x, expected, meta = fixtures.resnet20_files(tempfile.mkdtemp())
vm, _ = make_vm(lib, 15, 14)
lib.load(vm, cst.encode(), hv.encode()); lib.preprocess(vm)
f64p = C.POINTER(C.c_double)
lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size); lib.run(vm); lib.run(vm)  # lanes, then graph capture
l0 = lib.hevmx_param(vm, 6)
t = time.perf_counter(); lib.run(vm); dt = time.perf_counter() - t
print({k: os.environ.get(k) for k in ("HEVM_CHAIN", "HEVM_FUSE")}, "launches per run", lib.hevmx_param(vm, 6) - l0, "latency %.4f" % dt)
