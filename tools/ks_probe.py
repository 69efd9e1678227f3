"""Probe for ncu: a few rotate / mulcc / rescale ops at chosen levels inside a profiler range.
usage: python tools/ks_probe.py [levels...]   (default 13 4)"""
import ctypes as C
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm  # noqa: E402
from util import VM  # noqa: E402

levels = [int(x) for x in sys.argv[1:]] or [13, 4]
lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=6, npt=1)
for l in levels:
    for r in range(3):
        g.ct_write(r, g.random_ct(l, r))
    # warm
    g.exec(asm.ROTATE, 3, 0, 1); g.exec(asm.MULCC, 4, 0, 1)
    if l > 1:
        g.exec(asm.RESCALE, 5, 0)
    lib.hevmx_profiler_range(g.vm, 1)
    g.exec(asm.ROTATE, 3, 0, 2)
    g.exec(asm.MULCC, 4, 1, 2)
    if l > 1:
        g.exec(asm.RESCALE, 5, 2)
    g.exec(asm.ADDCC, 5, 0, 1)
    lib.hevmx_profiler_range(g.vm, 0)
print("probe done")
