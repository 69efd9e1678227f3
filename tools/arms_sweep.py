"""DaCapo vs PARS bootstrap placement over the waterline sweep 30..50 (BASELINE.json configs[3]) on the GPU: the programs of
tests/golden/resnet20_arms (restated reference pipelines, planned against profiled_B200_GPU.json) -- for every arm and
waterline the compiler's ESTIMATED latency next to the MEASURED run() latency and the rms against the plaintext model.
usage: python tools/arms_sweep.py [out.json]"""
import ctypes as C, json, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding
import fixtures
from util import make_vm

lib = _binding.bind(_binding.B200_LIB)
tmp = tempfile.mkdtemp()
f64p = C.POINTER(C.c_double)
allmeta = json.loads((fixtures.ARMS / "meta.json").read_text())
out = {"what": "encrypted ResNet-20 (nt = 2^14, N = 2^15, 14 x 60-bit primes) compiled by the restated reference pipelines: "
               "`pars` = ProactiveRescaling with the benchmark's 19 hand-placed bootstraps, `dacapo` = DaCapo planner; both against "
               "profiled_B200_GPU.json; estimated = sum of per-op table latencies (LatencyEstimator.cpp), measured = warm run() on one B200",
       "traced_ops": allmeta["traced"], "arms": {}}
vm, _ = make_vm(lib, 15, 14)
for name, m in sorted(allmeta["arms"].items()):
    if "failed" in m:
        out["arms"][name] = {"failed": m["failed"]}
        continue
    arm, W = name.split("_w")
    cst, hv, x, expected, meta = fixtures.resnet20_arm_files(tmp, arm, int(W))
    lib.load(vm, cst.encode(), hv.encode())
    lib.preprocess(vm)
    lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
    t0 = time.perf_counter(); lib.run(vm); first = time.perf_counter() - t0
    lat = []
    for _ in range(3):
        lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
        t0 = time.perf_counter(); lib.run(vm); lat.append(time.perf_counter() - t0)
    res = np.zeros(1 << 14)
    lib.decrypt_result(vm, 0, res.ctypes.data_as(f64p))
    r = res[:meta["n_out"]] * meta["post_scale"]
    out["arms"][name] = {"bootstraps": m["bootstraps"], "estimated_latency_s": round(m["estimated_latency_s"], 4),
                         "measured_run_s": round(float(np.median(lat)), 4), "first_run_s": round(first, 3),
                         "rms": float(np.sqrt(np.mean((r - expected) ** 2))), "argmax_ok": bool(np.argmax(r) == np.argmax(expected)),
                         "hevm_ops": m["hevm_ops"], "ops": m["ops"]}
    print(name, out["arms"][name], flush=True)
print(json.dumps(out))
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
