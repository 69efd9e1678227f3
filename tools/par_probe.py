"""How much do independent key switches overlap?  K independent rotations of one level-`lvl` ciphertext through
run() (graph replay, HEVM_STREAMS lanes) versus the solo latency.  usage: par_probe.py [level] [K]"""
import os, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench
from dacapo_b200 import _binding, hevm_asm as asm
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 13
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
vm = bench.make_vm(lib, tempfile.mkdtemp())
p = asm.Program(init_level=13)
x = p.arg(40, lvl)
outs = [p.new_ct() for _ in range(K)]
for i, o in enumerate(outs):
    p.rotate(o, x, 1 << (i % 13))
for o in outs[:1]:
    p.result(o, 40, lvl)
bench.load_program(lib, vm, p, tempfile.mkdtemp())
dat = np.linspace(-1, 1, 1 << 14)
import ctypes as C
lib.encrypt(vm, 0, dat.ctypes.data_as(C.POINTER(C.c_double)), dat.size)
lib.run(vm)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); lib.run(vm); ts.append(time.perf_counter() - t0)
print("level", lvl, "K", K, "streams", os.environ.get("HEVM_STREAMS"), "run us per rotation: %.1f" % (min(ts) * 1e6 / K))
