import ctypes as C, os, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench
from dacapo_b200 import _binding
lib = _binding.bind(_binding.B200_LIB)
vm = bench.make_vm(lib, tempfile.mkdtemp())
r = bench.resnet_mix(lib, vm, tempfile.mkdtemp(), cpu=False, reps=2)
print(os.environ.get("HEVM_STREAMS"), os.environ.get("HEVM_GRAPH"), "run", round(r["run_latency_s"], 4), "pre", round(r["preprocess_s"], 2))
