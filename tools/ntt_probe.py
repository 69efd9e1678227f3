"""Stand-alone forward NTT throughput: the two-launch path (pass A kernel, L2/HBM round trip, pass B kernel) vs the
single-pass cluster kernel (ntt_cluster.cuh).  usage: python tools/ntt_probe.py [limbs per prime]"""
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding  # noqa: E402
from util import VM  # noqa: E402

lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
g = VM(lib, 15, 14, keydir=tempfile.mkdtemp(), nct=2, npt=1, galois_steps=(1,))
for nb in [int(x) for x in sys.argv[1:]] or [182, 64, 16, 4]:
    for name, mode in (("two-launch", 0), ("cluster", 2)):
        ms = lib.hevmx_ntt_bench(g.vm, nb, 13, mode, 5)
        rate = nb * 13 / (ms * 1e-3)
        print(f"{nb:4d} limbs/prime  {name:10s}: {1e6 / rate:.3f} us per limb  ({rate * 2 * 262144 / 1e9:.0f} GB/s algorithmic)")
