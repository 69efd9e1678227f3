"""Throughput of batched launches (hevmx_exec_batch: n independent ciphertexts per kernel) vs op-by-op issue.
usage: python tools/batch_probe.py [levels...]"""
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm  # noqa: E402
from util import VM  # noqa: E402

levels = [int(x) for x in sys.argv[1:]] or [13, 4, 1]
BATCHES = [1, 8, 64, 256]
lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
nmax = max(BATCHES)
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=2 * nmax, npt=1)
for l in levels:
    for r in range(nmax):
        g.ct_write(r, g.random_ct(l, r % 4), 2.0 ** 40)
    for name, op, rhs in (("rotate", asm.ROTATE, 1), ("mulcc", asm.MULCC, None), ("rescale", asm.RESCALE, 0)):
        if op == asm.RESCALE and l < 2:
            continue
        row = {}
        for n in BATCHES:
            src, dst = list(range(n)), list(range(nmax, nmax + n))
            r = src if rhs is None else [rhs] * n
            reps = max(2, 256 // n)
            g.exec_batch(op, dst, src, r)
            lib.hevmx_timer(g.vm, 0)
            for _ in range(reps):
                g.exec_batch(op, dst, src, r, sync=False)
            ms = lib.hevmx_timer(g.vm, 1)
            row[n] = round(ms * 1e3 / (reps * n), 2)
        print(f"level {l:2d} {name:8s} us per op at batch {row}")
