"""Generates tests/golden/resnet20_arms/*: the encrypted ResNet-20 (nt = 2^14) compiled by the restated reference
pipelines -- `pars` (dacapo_b200.earth: ProactiveRescaling + EarlyModswitch, the benchmark's 19 hand-placed bootstraps
honoured element-wise) and `dacapo` (dacapo_b200.dacapo: automatic placement against profiled_B200_GPU.json) -- for the
waterlines 30..50 of BASELINE.json configs[3].

Runs ONLY in the build container (imports the reference's Python from /root/reference, like make_resnet_fixture.py).
Input, expected logits and post-processing are those of tests/golden/resnet20 (same benchmark, same packing).

    python tests/golden/make_resnet_pars_dacapo.py
"""
import json
import lzma
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "python" / "poly"))

from dacapo_b200 import dacapo, earth, frontend  # noqa: E402

sys.modules["hecate"] = frontend


def trace(live_bootstraps):
    bench = REF / "examples" / "benchmarks" / "ResNet.py"
    src = bench.read_text()
    assert '"nt" : 2**16,' in src
    src = src.replace('"nt" : 2**16,', '"nt" : 2**14,')
    frontend.reset()
    frontend.LIVE_LIST_BOOTSTRAPS = live_bootstraps
    g = {"__name__": "__main__", "__file__": str(bench)}
    try:
        exec(compile(src, str(bench), "exec"), g)
    finally:
        frontend.LIVE_LIST_BOOTSTRAPS = False
    return g["modName"]


def main():
    out = HERE / "resnet20_arms"
    out.mkdir(exist_ok=True)
    prof = json.loads((REPO / "profiled_B200_GPU.json").read_text())
    g_auto, g_manual = trace(False), trace(True)
    meta = {"source": "examples/benchmarks/ResNet.py (nt=2^14) traced with dacapo_b200.frontend", "cost_profile": "profiled_B200_GPU.json",
            "traced": earth.op_counts(earth.from_graph(g_manual)), "arms": {}}
    cst = None
    for W in (40, 30, 35, 45, 50):
        P = earth.Params.from_profile(prof, waterline=W, output_val=10)
        for arm in ("pars", "dacapo"):
            t0 = time.time()
            try:
                if arm == "pars":
                    prog, f = earth.compile_pars(g_manual, P)
                    rep = {"bootstraps": earth.op_counts(f).get("boot", 0), "estimated_latency_s": earth.latency(f, P) / 1e6}
                else:
                    prog, f, rep = dacapo.compile_dacapo(g_auto, P)
            except earth.PassFailed as e:
                meta["arms"][f"{arm}_w{W}"] = {"failed": str(e)}
                print(arm, W, "FAILED", e)
                continue
            raw = prog.cst_bytes()
            if cst is None:
                cst = raw
                (out / "consts.cst.xz").write_bytes(lzma.compress(raw, preset=6))
            if raw != cst:  # another constant order: this arm gets its own pool
                (out / f"{arm}_w{W}.cst.xz").write_bytes(lzma.compress(raw, preset=6))
            (out / f"{arm}_w{W}.hevm").write_bytes(prog.hevm_bytes())
            meta["arms"][f"{arm}_w{W}"] = {**rep, "ops": earth.op_counts(f), "hevm_ops": len(prog.ops), "ct_registers": prog.num_ct,
                                           "pt_registers": prog.num_pt, "own_constant_pool": raw != cst, "compile_s": round(time.time() - t0, 1)}
            print(arm, W, meta["arms"][f"{arm}_w{W}"], flush=True)
    (out / "meta.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
