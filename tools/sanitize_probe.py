"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): one of every kernel family at a small level --
rotate, multiply + relinearise, rescale (separate launches and the batched single-launch form), bootstrap, the cluster
NTT.  usage: compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm  # noqa: E402
from util import VM  # noqa: E402

lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, 15, 5, keydir=tempfile.mkdtemp(), nct=10, npt=2, galois_steps=(1, -2), env={"HEVM_STREAMS": 2})
lvl = 3
for r in range(4):
    g.ct_write(r, g.random_ct(lvl, r), 2.0 ** 40)
g.exec(asm.ROTATE, 4, 0, 1)
g.exec(asm.MULCC, 5, 0, 1)
g.exec(asm.RESCALE, 6, 5)
g.exec(asm.ADDCC, 7, 0, 1)
g.exec_batch(asm.ROTATE, [4, 5], [0, 1], [1, -2])
g.exec_batch(asm.MULCC, [6, 7], [0, 1], [2, 3])
g.exec_batch(asm.RESCALE, [8, 9], [6, 7], [0, 0])
x = np.random.default_rng(0).uniform(-1, 1, g.N // 2)
g.encode(0, x, 2, 40)
g.encrypt_pt(0, 0, counter=1)
g.exec(asm.BOOTSTRAP, 1, 0, 4)
f = np.random.default_rng(1).integers(0, g.primes[0], size=(3, g.N), dtype=np.uint64)
a, b = g.ntt(f, 0), g.ntt(f, 0, inverse=2)
assert np.array_equal(a, b)
print("sanitize probe done; decrypt error", float(np.max(np.abs(g.decrypt_decode(1, 1) - x))))

# the C ABI flow: asynchronous encrypt() on several lanes, run() three times (lanes / graph capture / replay), batched decrypt_result()
import ctypes as C
p = asm.Program(init_level=4)
xs = [p.arg(50, 3), p.arg(50, 3), p.arg(40, 3)]  # 50 + 50 - 60 (rescale) = 40 bits, the scale of the rotated operand
t, u = p.new_ct(), p.new_ct()
p.emit(asm.MULCC, t, xs[0], xs[1])
p.emit(asm.RESCALE, t, t)
p.rotate(u, xs[2], 1)
p.emit(asm.MODSWITCH, u, u, 1)
p.emit(asm.ADDCC, t, t, u)
p.emit(asm.BOOTSTRAP, u, t, 3)
p.result(t, 40, 2)
p.result(u, 40, 3)
d = tempfile.mkdtemp()
p.save(d + "/s.cst", d + "/s.hevm")
lib.load(g.vm, (d + "/s.cst").encode(), (d + "/s.hevm").encode())
lib.preprocess(g.vm)
f64p = C.POINTER(C.c_double)
vals = np.random.default_rng(2).uniform(-1, 1, (3, g.N // 2))
res = np.zeros((2, g.N // 2))
for rep in range(3):
    for i in range(3):
        lib.encrypt(g.vm, i, vals[i].ctypes.data_as(f64p), g.N // 2)
    lib.run(g.vm)
    for i in range(2):
        lib.decrypt_result(g.vm, i, res[i].ctypes.data_as(f64p))
exp = vals[0] * vals[1] + np.roll(vals[2], -1)
print("C ABI flow done; errors", float(np.max(np.abs(res[0] - exp))), float(np.max(np.abs(res[1] - exp))))
