"""Measured cost profile (schema of EarthDialect.cpp:130-180) for another ring geometry.
usage: python tools/emit_profile.py <logN> <num_primes> <out.json> [reps]"""
import sys, tempfile
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, profile
from util import VM

logn, npr, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, logn, npr, keydir=tempfile.mkdtemp(), nct=18, npt=2)
table, prof = profile.emit_profile(out, lib, g.vm, reps=reps, kernel_profile_level=None)
print(out, "rotate us:", prof["latencyTableExact"]["earth.rotate_single"][::4])
