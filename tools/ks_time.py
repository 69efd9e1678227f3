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
import os
lib = _binding.bind(os.environ.get('HEVM_LIB', _binding.B200_LIB))
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=18, npt=2)
tab = profile.measure_op_table(lib, g.vm, levels=levels, reps=30, kernel_profile_level=levels[0])
kern = tab.pop('_kernels_rotate_l%d' % levels[0], {})
print('kernels @%d:' % levels[0], {k: round(v['ms'] * 1e3 / v['launches'], 1) for k, v in kern.items()})
for op in ("rotate", "mulcc", "rescale", "addcc", "bootstrap"):
    print(op, {l: round(v, 1) for l, v in tab[op].items()})
