"""Per-kernel-class time of one in-order run() of the ResNet-20 fixture (hevmx_profile: CUDA events around every launch,
lane 0 only) next to the latency of the scheduled graph replay: how much device time the program holds per class."""
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
t = time.perf_counter(); lib.run(vm); dt = time.perf_counter() - t
print("graph replay latency %.4f s" % dt)
lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
lib.hevmx_profile(vm, 1)
t = time.perf_counter(); lib.run(vm); dts = time.perf_counter() - t
lib.hevmx_profile(vm, 0)
tot, rows = 0.0, []
for cls in range(64):
    ms, cnt = C.c_double(), C.c_int64()
    name = lib.hevmx_profile_read(vm, cls, C.byref(ms), C.byref(cnt))
    if not name:
        break
    if cnt.value:
        rows.append((ms.value, cnt.value, name.decode()))
        tot += ms.value
print("in-order profiled run %.4f s wall, %.2f ms summed kernel time" % (dts, tot))
for ms, cnt, name in sorted(rows, reverse=True):
    print("%-28s %7d launches %9.3f ms  %6.2f us avg  %5.1f %%" % (name, cnt, ms, 1e3 * ms / cnt, 100 * ms / tot))
