"""Generates tests/golden/params_n15.json from the independent Python big-integer
reference (tests/pyref.py).  The reference repository holds no golden vectors for this
path, so this pins the full-size parameter set (SEAL 4.0
CoeffModulus::Create(2^15, {60 x 14}) + minimal primitive 2N-th roots) for the
oracle and the CUDA library.  Run: python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pyref  # noqa: E402

N = 1 << 15
primes = pyref.seal_primes(N, 60, 14)
roots = [pyref.minimal_root(N, q) for q in primes]
Path(__file__).with_name("params_n15.json").write_text(json.dumps({"N": N, "primes": primes, "roots": roots}, indent=1))
print("wrote params_n15.json")
