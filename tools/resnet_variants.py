"""BUILD CONTAINER ONLY (imports the reference's benchmark from /root/reference): compile the traced ResNet-20 (nt = 2^14)
against alternative cost tables -- the lone-op latency table of profiled_B200_GPU.json, a throughput table (per-op time
when many independent ops are in flight: fitted to bench.py's batch sweep) and blends of the two -- and write one .hevm per
table into tests/golden/resnet20_variants/ (same constant pool as tests/golden/resnet20).  The GPU side
(tools/resnet_variants_run.py) measures run() and rms of every variant."""
import json
import lzma
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "python" / "poly"))
from dacapo_b200 import compiler, frontend  # noqa: E402

sys.modules["hecate"] = frontend
OUT = REPO / "tests" / "golden" / "resnet20_variants"


def throughput_table(lone, ks0=15.0, boot_a=25.0, boot_b=6.0):
    L = range(1, 14)
    ks = lambda l: l * l + 3 * l
    t = {k: list(v) for k, v in lone.items()}
    t["earth.rotate_single"] = [ks0 + 0.485 * ks(l) for l in L]   # bench.py batch sweep (64 ciphertexts per launch)
    t["earth.mul_double"] = [ks0 - 9.8 + 0.544 * ks(l) for l in L]
    t["earth.rescale_single"] = [0.6 + 1.0 * l for l in L]
    for k in ("earth.add_single", "earth.add_double", "earth.negate_single", "earth.modswitch_single"):
        t[k] = [0.5 + 0.3 * l for l in L]
    t["earth.mul_single"] = [0.5 + 0.4 * l for l in L]
    t["earth.bootstrap_single"] = [boot_a + boot_b * l for l in L]  # ~ (4 l + 10) limb transforms + FFTs, many small launches
    return t


def blend(a, b, w):
    return {k: [w * x + (1 - w) * y for x, y in zip(a[k], b[k])] for k in a}


def main():
    bench = REF / "examples" / "benchmarks" / "ResNet.py"
    src = bench.read_text().replace('"nt" : 2**16,', '"nt" : 2**14,')
    g = {"__name__": "__main__", "__file__": str(bench)}
    frontend.reset()
    exec(compile(src, str(bench), "exec"), g)
    graph = g["modName"]
    lone = json.loads((REPO / "profiled_B200_GPU.json").read_text())["latencyTableExact"]
    tp = throughput_table(lone)
    raw = lzma.open(REPO / "tests" / "golden" / "resnet20" / "resnet20.cst.xz").read()
    OUT.mkdir(exist_ok=True)
    meta = {}
    variants = [("lone", lone), ("blend50", blend(lone, tp, 0.5)), ("blend25", blend(lone, tp, 0.25)), ("throughput", tp),
                ("tp_boot15_4", throughput_table(lone, 15.0, 15.0, 4.0)), ("tp_boot10_3", throughput_table(lone, 15.0, 10.0, 3.0)),
                ("tp_ks10", throughput_table(lone, 10.0, 25.0, 6.0)), ("tp_ks10_boot10_3", throughput_table(lone, 10.0, 10.0, 3.0)),
                ("tp_boot40_6", throughput_table(lone, 15.0, 40.0, 6.0))]
    for name, table in variants:
        t0 = time.time()
        prog, c = compiler.compile_graph(graph, compiler.Options(logN=15, num_primes=14, waterline=40, cost_table=table))
        same = prog.cst_bytes() == raw
        print(name, "compiled in %.0fs" % (time.time() - t0), c.stats, "ops", len(prog.ops), "same constant pool:", same, flush=True)
        if not same:
            continue
        (OUT / f"{name}.hevm").write_bytes(prog.hevm_bytes())
        meta[name] = {"lowered_ops": c.stats, "hevm_ops": len(prog.ops), "table": table}
    (OUT / "meta.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
