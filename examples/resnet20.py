"""`hc-test`-style run of the encrypted ResNet-20 on the B200 backend, mirroring the reference's
examples/tests/ResNet.py:85-118 (load -> setInput -> run (timed) -> getOutput -> rms -> printer):

    python examples/resnet20.py b200c 40 B200 GPU

The program / constants are the committed fixture (tests/golden/resnet20, traced from the reference's
examples/benchmarks/ResNet.py at nt = 2^14 and compiled by dacapo_b200.compiler); the expected logits come from the
plaintext torch model with the reference's weights.  Reference README.md:186-187: latency 53.7 s, rms 9.5e-4 on CPU.
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dacapo_b200 as hc  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import fixtures  # noqa: E402  (committed test fixture: tests/fixtures.py)

if __name__ == "__main__":
    hc.setLibnHW(sys.argv)
    cst, hv, x, expected, meta = fixtures.resnet20_files(tempfile.mkdtemp())
    hevm = hc.HEVM()
    hevm.load(cst, hv)
    hevm.setInput(0, x)
    first = time.perf_counter_ns()
    hevm.run()                      # first run of the process: issued on the lanes (what the reference's hc-test times)
    first = time.perf_counter_ns() - first
    hevm.setInput(0, x)
    hevm.run()                      # second run: captured into a CUDA graph
    hevm.setInput(0, x)
    timer = time.perf_counter_ns()
    hevm.run()
    timer = time.perf_counter_ns() - timer
    res = hevm.getOutput()[0][:meta["n_out"]] * meta["post_scale"]
    err = res - expected
    rms = np.sqrt(np.sum(err * err) / res.shape[-1])
    print("logits (encrypted):", np.round(res, 4))
    print("logits (plaintext):", np.round(expected, 4))
    print("first run of the process (s):", first / 1e9)
    hevm.printer(timer / 1e9, rms)
