"""Generates tests/golden/bench/<name>/* for the reference's small benchmarks (examples/benchmarks/*.py):
SobelFilter, HarrisCornerDetection, LinearRegression, PolynomialRegression, Multivariate, MLP.

Runs ONLY in the build container (needs /root/reference).  Every benchmark script is executed unmodified against
`dacapo_b200.frontend` registered as `hecate`; the recorded graph is compiled by dacapo_b200.compiler (waterline 40,
N = 2^15, 14 primes, bootstrap levels from profiled_B200_GPU.json) and stored with seeded inputs and the expected
outputs = the plaintext evaluation of the same graph (frontend.evaluate).

    python tests/golden/make_bench_fixtures.py
"""
import json
import lzma
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "python" / "poly"))

from dacapo_b200 import compiler, frontend  # noqa: E402

sys.modules["hecate"] = frontend

# input length and value range per benchmark (examples/tests/<name>.py: smooth 64x64 images, 4096 samples in (-1,1),
# 784 pixels + 16 zeros for the MLP)
BENCHES = {
    "SobelFilter": (4096, 0.4, 0.6),
    "HarrisCornerDetection": (4096, 0.4, 0.6),
    "LinearRegression": (4096, -1.0, 1.0),
    "PolynomialRegression": (4096, -1.0, 1.0),
    "Multivariate": (4096, -1.0, 1.0),
    "MLP": (784, 0.0, 1.0),
}


def main():
    cost = json.loads((REPO / "profiled_B200_GPU.json").read_text())["latencyTableExact"]
    nslots = 1 << 14
    for name, (n, lo, hi) in BENCHES.items():
        bench = REF / "examples" / "benchmarks" / f"{name}.py"
        g = {"__name__": "__main__", "__file__": str(bench)}
        frontend.reset()
        exec(compile(bench.read_text(), str(bench), "exec"), g)
        graph = g["modName"]
        prog, c = compiler.compile_graph(graph, compiler.Options(logN=15, num_primes=14, waterline=40, cost_table=cost))
        rng = np.random.default_rng(abs(hash(name)) % (1 << 31) if False else sum(map(ord, name)))
        inputs = rng.uniform(lo, hi, size=(graph.n_inputs, n))
        expected = np.stack(frontend.evaluate(graph, inputs, nslots))
        out = HERE / "bench" / name
        out.mkdir(parents=True, exist_ok=True)
        (out / "prog.hevm").write_bytes(prog.hevm_bytes())
        (out / "prog.cst.xz").write_bytes(lzma.compress(prog.cst_bytes(), preset=6))
        np.savez_compressed(out / "io.npz", inputs=inputs, expected=expected.astype(np.float64))
        meta = {"source": f"examples/benchmarks/{name}.py", "lowered_ops": c.stats, "n_inputs": graph.n_inputs,
                "n_outputs": len(graph.outputs), "input_len": n, "max_abs_expected": float(np.max(np.abs(expected)))}
        (out / "meta.json").write_text(json.dumps(meta, indent=1))
        print(name, c.stats, "max|expected| %.3g" % meta["max_abs_expected"])


if __name__ == "__main__":
    main()
