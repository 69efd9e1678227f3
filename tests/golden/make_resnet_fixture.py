"""Generates tests/golden/resnet20/* : the REAL encrypted ResNet-20 workload of BASELINE.json configs[0].

Runs ONLY in the build container (it imports the reference's Python from /root/reference, which does not
exist on the GPU box): the reference benchmark script examples/benchmarks/ResNet.py is executed unmodified
(except for the documented `nt = 2**14` edit, README.md:170-171) against `dacapo_b200.frontend` registered
as the `hecate` module, which records the op graph the reference would hand to libHecateFrontend.so.
The graph is compiled with dacapo_b200.compiler (waterline 40, N = 2^15, 14 primes) and stored with
 - resnet20.hevm           the program (reference container format)
 - resnet20.cst.xz         the 4 000+ plaintext constants (525 MB raw, ~4 MB compressed: <= 17 distinct values per row)
 - input.npy               the packed, seeded synthetic input (CIFAR-10 cannot be downloaded offline)
 - expected.npy            the plaintext torch model's logits for that input (the reference test's `process`)
 - meta.json               post-processing (x32, first 10 slots: examples/tests/ResNet.py:76-82) and statistics

    python tests/golden/make_resnet_fixture.py
"""
import json
import lzma
import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "python" / "poly"))

from dacapo_b200 import compiler, frontend  # noqa: E402

sys.modules["hecate"] = frontend  # the reference scripts do `import hecate as hc`


def make(nt_bits, logN, out, cost, do_sweep):
    import torch
    out.mkdir(exist_ok=True)
    bench = REF / "examples" / "benchmarks" / "ResNet.py"
    src = bench.read_text()
    assert '"nt" : 2**16,' in src
    src = src.replace('"nt" : 2**16,', f'"nt" : 2**{nt_bits},')
    g = {"__name__": "__main__", "__file__": str(bench)}
    t0 = time.time()
    frontend.reset()
    exec(compile(src, str(bench), "exec"), g)
    graph = g["modName"]  # frontend.save() returns the recorded Graph
    print(f"traced {len(graph.nodes)} nodes, {len(graph.consts)} constants in {time.time() - t0:.1f}s")

    # bootstrap levels are chosen against the measured B200 cost profile of this ring size
    prog, c = compiler.compile_graph(graph, compiler.Options(logN=logN, num_primes=14, waterline=40, cost_table=cost))
    print("lowered ops:", c.stats, "ct regs", prog.num_ct, "pt regs", prog.num_pt, "pool", len(prog.constants))
    (out / "resnet20.hevm").write_bytes(prog.hevm_bytes())
    raw = prog.cst_bytes()
    (out / "resnet20.cst.xz").write_bytes(lzma.compress(raw, preset=6))
    print("cst raw MB", len(raw) / 1e6, "xz MB", (out / "resnet20.cst.xz").stat().st_size / 1e6)

    # waterline sweep (BASELINE.json configs[3]): same graph, same constant pool, other scale-management waterlines
    sweep = {}
    for W in (30, 35, 45, 50) if do_sweep else ():
        pw, cw = compiler.compile_graph(graph, compiler.Options(logN=logN, num_primes=14, waterline=W, cost_table=cost))
        if pw.cst_bytes() != raw:
            print(f"waterline {W}: constant pool differs, skipped")
            continue
        (out / f"resnet20_w{W}.hevm").write_bytes(pw.hevm_bytes())
        sweep[str(W)] = {"lowered_ops": cw.stats, "hevm_ops": len(pw.ops)}
        print(f"waterline {W}:", cw.stats)

    # seeded synthetic input, plaintext reference and packing -- with the reference's own code
    model = g["getModel"]()
    torch.manual_seed(1)
    x = torch.randn(1, 3, 32, 32).clamp(-2.0, 2.0)
    with torch.no_grad():
        logits = model(x).cpu().numpy()[0].astype(np.float64)
    shapes = {"nt": 2 ** nt_bits, "bb": 32, "ko": 1, "ho": 32, "wo": 32}
    conv1_shapes = g["CascadeConv"](shapes, model.module.conv1)
    close = g["shapeClosure"](**conv1_shapes)
    packed = np.asarray(close["MPP"](x)[0], dtype=np.float64).ravel()
    if do_sweep:
        np.save(out / "input.npy", packed)
    else:
        np.savez_compressed(out / "input.npz", packed=packed)
    np.save(out / "expected.npy", logits)
    meta = {"post_scale": 32.0, "n_out": 10, "lowered_ops": c.stats, "traced_nodes": len(graph.nodes),
            "ct_registers": prog.num_ct, "pt_registers": prog.num_pt, "hevm_ops": len(prog.ops), "waterline": 40,
            "waterline_sweep": sweep,
            "logN": logN, "slots": 2 ** nt_bits,
            "source": f"examples/benchmarks/ResNet.py (nt=2^{nt_bits}) traced with dacapo_b200.frontend, compiled with dacapo_b200.compiler",
            "weights": "examples/data/resnet20.silu.model", "input": "torch.manual_seed(1); randn(1,3,32,32).clamp(-2,2)"}
    (out / "meta.json").write_text(json.dumps(meta, indent=1))
    print("logits", logits)


def main():
    # planner cost = per-op time with many independent ops in flight (profile.measure_throughput_table): the backend overlaps
    # independent ops on 16 lanes, so the sum a planner minimises must be built from those, not from lone-op latencies
    prof15 = json.loads((REPO / "profiled_B200_GPU.json").read_text())
    cost15 = prof15.get("latencyTableThroughput") or prof15["latencyTableExact"]
    make(14, 15, HERE / "resnet20", cost15, True)
    # BASELINE.json configs[2]: nt = 2^16 slots (the benchmark's own default) => N = 2^17; same 14 x 60-bit chain
    cost17 = json.loads((REPO / "profiles" / "profiled_B200_GPU_N17_L30.json").read_text())["latencyTableExact"]
    make(16, 17, HERE / "resnet20_nt16", {k: v[:13] for k, v in cost17.items()}, False)


if __name__ == "__main__":
    main()
