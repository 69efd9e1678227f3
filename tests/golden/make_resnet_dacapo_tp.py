"""Adds the arm `dacapotp_w40` to tests/golden/resnet20_arms: the restated DaCapo planner (dacapo_b200.dacapo) run against
`latencyTableThroughput` of profiled_B200_GPU.json -- per-op time with 32 independent ops in flight -- instead of the
lone-op `latencyTable` the `dacapo` arm uses.  Build container only (imports /root/reference like make_resnet_pars_dacapo.py).

    python tests/golden/make_resnet_dacapo_tp.py
"""
import json
import lzma
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_resnet_pars_dacapo as base  # noqa: E402  (registers dacapo_b200.frontend as `hecate`)
from dacapo_b200 import dacapo, earth  # noqa: E402


def main():
    out = HERE / "resnet20_arms"
    prof = json.loads((base.REPO / "profiled_B200_GPU.json").read_text())
    prof_tp = dict(prof, latencyTableExact=prof["latencyTableThroughput"])
    P = earth.Params.from_profile(prof_tp, waterline=40, output_val=10, exact=True)
    g_auto = base.trace(False)
    t0 = time.time()
    prog, f, rep = dacapo.compile_dacapo(g_auto, P)
    raw = prog.cst_bytes()
    shared = lzma.open(out / "consts.cst.xz").read()
    own = raw != shared
    if own:
        (out / "dacapotp_w40.cst.xz").write_bytes(lzma.compress(raw, preset=6))
    (out / "dacapotp_w40.hevm").write_bytes(prog.hevm_bytes())
    meta = json.loads((out / "meta.json").read_text())
    # the estimate of this arm is in under-load microseconds; the lone-op estimate of the same program for comparison
    P_lone = earth.Params.from_profile(prof, waterline=40, output_val=10)
    meta["arms"]["dacapotp_w40"] = {**rep, "estimated_latency_lone_table_s": earth.latency(f, P_lone) / 1e6, "ops": earth.op_counts(f),
                                    "hevm_ops": len(prog.ops), "ct_registers": prog.num_ct, "pt_registers": prog.num_pt,
                                    "own_constant_pool": own, "compile_s": round(time.time() - t0, 1),
                                    "cost_table": "latencyTableThroughput (profile.measure_throughput_table)"}
    (out / "meta.json").write_text(json.dumps(meta, indent=1))
    print(meta["arms"]["dacapotp_w40"])


if __name__ == "__main__":
    main()
