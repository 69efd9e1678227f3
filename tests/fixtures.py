"""TEST / BENCH INFRASTRUCTURE: access to the committed ResNet-20 fixture (tests/golden/resnet20, made by
tests/golden/make_resnet_fixture.py) and the location of the CPU oracle library.  Not part of the product package."""
import json
import lzma
import tempfile
from pathlib import Path

import numpy as np

FIXTURE = Path(__file__).resolve().parent / "golden" / "resnet20"
ORACLE_LIB = Path(__file__).resolve().parent.parent / "oracle" / "libORACLE_HEVM.so"  # the checker, never the product


def resnet20_files(tmpdir=None, variant=""):
    """Decompress the constants next to a copy of the program; returns (cst_path, hevm_path, input, expected, meta).

    The directory layout mimics the reference's `optimized/<compiler>/<bench>.<waterline>._hecate_<bench>.hevm`
    so that HEVM.printer (runner.py:256-271) can parse it."""
    fixture = FIXTURE.parent / ("resnet20" + variant)  # variant "_nt16": nt = 2^16 slots, N = 2^17
    tmp = Path(tmpdir or tempfile.mkdtemp(prefix="resnet20_"))
    cst = tmp / "traced" / "_hecate_ResNet.cst"
    hv = tmp / "optimized" / "b200c" / "ResNet.40._hecate_ResNet.hevm"
    cst.parent.mkdir(parents=True, exist_ok=True)
    hv.parent.mkdir(parents=True, exist_ok=True)
    if not cst.is_file():
        with lzma.open(fixture / "resnet20.cst.xz") as f, open(cst, "wb") as o:
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                o.write(b)
    hv.write_bytes((fixture / "resnet20.hevm").read_bytes())
    meta = json.loads((fixture / "meta.json").read_text())
    x = np.load(fixture / "input.npy") if (fixture / "input.npy").is_file() else np.load(fixture / "input.npz")["packed"]
    return str(cst), str(hv), x, np.load(fixture / "expected.npy"), meta


ARMS = FIXTURE.parent / "resnet20_arms"


def resnet20_arm_files(tmpdir, arm="dacapo", waterline=40):
    """The ResNet-20 program compiled by the restated reference pipelines (tests/golden/make_resnet_pars_dacapo.py):
    arm = "pars" (hand-placed bootstraps, ProactiveRescaling) | "dacapo" (automatic placement).  Input / expected logits
    are those of the resnet20 fixture.  Returns (cst_path, hevm_path, input, expected, meta of the arm)."""
    tmp = Path(tmpdir)
    name = f"{arm}_w{waterline}"
    allmeta = json.loads((ARMS / "meta.json").read_text())
    meta = allmeta["arms"][name]
    src = ARMS / (f"{name}.cst.xz" if meta.get("own_constant_pool") else "consts.cst.xz")
    cst = tmp / src.name[:-3]
    hv = tmp / f"{name}.hevm"
    if not cst.is_file():
        with lzma.open(src) as f, open(cst, "wb") as o:
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                o.write(b)
    hv.write_bytes((ARMS / f"{name}.hevm").read_bytes())
    base = json.loads((FIXTURE / "meta.json").read_text())
    return str(cst), str(hv), np.load(FIXTURE / "input.npy"), np.load(FIXTURE / "expected.npy"), {**base, **meta}
