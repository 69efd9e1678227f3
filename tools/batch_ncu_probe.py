"""One batched key switch (hevmx_exec_batch, single-launch dataflow kernel) inside a cudaProfilerStart/Stop range.
usage: ncu --set full --profile-from-start off -o out python tools/batch_ncu_probe.py [level] [batch] [rotate|mulcc]"""
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm  # noqa: E402
from util import VM  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
what = sys.argv[3] if len(sys.argv) > 3 else "rotate"
lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=2 * n, npt=1)
for r in range(n):
    g.ct_write(r, g.random_ct(level, r % 4), 2.0 ** 40)
src, dst = list(range(n)), list(range(n, 2 * n))
op, rhs = (asm.ROTATE, [1] * n) if what == "rotate" else (asm.MULCC, src)
for _ in range(3):
    g.exec_batch(op, dst, src, rhs)
lib.hevmx_timer(g.vm, 0)
for _ in range(8):
    g.exec_batch(op, dst, src, rhs, sync=False)
ms = lib.hevmx_timer(g.vm, 1)
print(f"{what} level {level} batch {n}: {ms * 1e3 / (8 * n):.2f} us per op")
lib.hevmx_profiler_range(g.vm, 1)
g.exec_batch(op, dst, src, rhs)
lib.hevmx_profiler_range(g.vm, 0)
