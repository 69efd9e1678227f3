"""Builds libB200_HEVM.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libB200_HEVM.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--fmad=true"]


def needs_build():
    if not OUT.is_file():
        return True
    t = OUT.stat().st_mtime
    return any(f.stat().st_mtime > t for f in CSRC.iterdir())


def build(force=False, verbose=False, out=OUT, defs=()):
    if not force and out == OUT and not needs_build():
        return OUT
    objs = []
    for src in ("kernels.cu", "vm.cu"):
        obj = CSRC / (src + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defs], "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    # (the driver API -- cuTensorMapEncodeTiled for the TMA tensor map of the cluster NTT -- is reached through
    #  cudaGetDriverEntryPoint, so libcuda is not a link-time dependency and the library still loads on a CPU-only box)
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
