"""One wrapper bootstrap (decrypt + re-encrypt on device) inside a profiler range, for `ncu --profile-from-start off`."""
import sys, tempfile
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm
from util import VM
src, tgt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 3)
lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=4, npt=1)
g.ct_write(0, g.random_ct(src, 1))
g.exec(asm.BOOTSTRAP, 1, 0, tgt)
lib.hevmx_profiler_range(g.vm, 1)
g.exec(asm.BOOTSTRAP, 1, 0, tgt)
lib.hevmx_profiler_range(g.vm, 0)
print("done")
