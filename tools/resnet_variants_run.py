"""Run every ResNet-20 variant of tests/golden/resnet20_variants (tools/resnet_variants.py) on the GPU: warm run() latency,
rms against the plaintext logits, op counts."""
import ctypes as C, json, lzma, os, sys, tempfile, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding
import fixtures
from util import make_vm
lib = _binding.bind(os.environ.get("HEVM_LIB", _binding.B200_LIB))
tmp = tempfile.mkdtemp()
cst, hv0, x, expected, meta0 = fixtures.resnet20_files(tmp)
VAR = REPO / "tests" / "golden" / "resnet20_variants"
meta = json.loads((VAR / "meta.json").read_text())
vm, _ = make_vm(lib, 15, 14)
f64p = C.POINTER(C.c_double)
out = {}
for name in meta:
    hv = VAR / f"{name}.hevm"
    lib.load(vm, cst.encode(), str(hv).encode()); lib.preprocess(vm)
    lat = []
    for i in range(5):
        lib.encrypt(vm, 0, x.ctypes.data_as(f64p), x.size)
        t = time.perf_counter(); lib.run(vm); lat.append(time.perf_counter() - t)
    res = np.zeros(1 << 14)
    lib.decrypt_result(vm, 0, res.ctypes.data_as(f64p))
    r = res[:meta0["n_out"]] * meta0["post_scale"]
    rms = float(np.sqrt(np.sum((r - expected) ** 2) / r.shape[-1]))
    ops = meta[name]["lowered_ops"]
    out[name] = {"run_s": min(lat[2:]), "first_s": lat[0], "rms": rms, "argmax_ok": bool(np.argmax(r) == np.argmax(expected)),
                 "bootstrap": ops.get("bootstrap"), "rotate": ops.get("rotate"), "rescale": ops.get("rescale"), "hevm_ops": meta[name]["hevm_ops"]}
    print(name, json.dumps(out[name]), flush=True)
(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "resnet_variants.json").write_text(json.dumps(out, indent=1))
