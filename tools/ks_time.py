"""Quick CUDA-event timing of rotate / mulcc / rescale at a few levels (no ncu)."""
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm, profile  # noqa: E402
from util import VM  # noqa: E402

levels = [int(x) for x in sys.argv[1:]] or [13, 8, 4, 2]
lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=18, npt=2)
tab = profile.measure_op_table(lib, g.vm, levels=levels, reps=30)
for op in ("rotate", "mulcc", "rescale", "addcc", "bootstrap"):
    print(op, {l: round(v, 1) for l, v in tab[op].items()})
