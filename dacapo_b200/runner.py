"""Host-side driver: same class / function names, argument meaning and error behaviour
as the reference's Python driver (reference: python/hecate/hecate/runner.py:123-271),
with the `B200 GPU` target added to its library/hardware allow-list (runner.py:126-130).

    hc.setLibnHW(sys.argv)          # argv[3:5] = "B200" "GPU"
    vm = hc.HEVM()                  # keys under ~/.hevm/b200
    vm.load(cst, hevm); vm.setInput(0, x); vm.run(); out = vm.getOutput()

The product path is libB200_HEVM.so only.  There is NO CPU fallback: if the CUDA
library is missing, or no GPU is visible, construction fails loudly.
"""
import ctypes
import os
import re
from pathlib import Path

import numpy as np

from . import _binding

# library -> supported hardware (reference runner.py:126-130 lists {"SEAL": ["CPU"]})
_LIBNHW = {"B200": ["GPU"]}
run_library = "B200"
run_hardware = "GPU"
lw = None


def _lib_path(library):
    if library == "B200":
        return _binding.B200_LIB
    raise ValueError(f"unsupported library {library}")


def reinit_lw():
    """(Re)load the shared library for the selected target (reference runner.py:73-117)."""
    global lw
    lw = _binding.bind(_lib_path(run_library))
    return lw


def setLibnHW(argv=None):
    """Select library/hardware from argv[3:5] in either order (reference runner.py:123-171)."""
    global run_library, run_hardware
    hw2lib = {}
    for lib, hws in _LIBNHW.items():
        for hw in hws:
            hw2lib.setdefault(hw, []).append(lib)
    argv = list(argv or [])
    if len(argv) >= 4:
        a3 = argv[3].upper()
        a4 = argv[4].upper() if len(argv) >= 5 else None
        if a3 in _LIBNHW:
            run_library = a3
            if a4 is not None and a4 not in _LIBNHW[a3]:
                print("Not supported", a4)
                print("Supported hardware :", _LIBNHW[a3])
                raise SystemExit(1)
            run_hardware = a4 or _LIBNHW[a3][0]
        elif a3 in hw2lib:
            run_hardware = a3
            if a4 is not None and a4 not in hw2lib[a3]:
                print("Not supported", a4)
                print("Supported library :", hw2lib[a3])
                raise SystemExit(1)
            run_library = a4 or hw2lib[a3][0]
        else:
            print("Not supported", argv[3])
            print("Supported library :", list(_LIBNHW))
            print("Supported hardware :", list(hw2lib))
            raise SystemExit(1)
    else:
        run_library = next(iter(_LIBNHW))
        run_hardware = _LIBNHW[run_library][0]


class HEVM:
    """Mirror of the reference `HEVM` class (runner.py:174-271)."""

    def __init__(self, path=None, option="full", lib=None, seal_dir=None):
        """`seal_dir`: a SEAL 4.0 key directory (parm/pub/sec/relin/gal.seal as written by SEAL_HEVM.cpp:56-88) whose keys
        replace the seed-derived ones (dacapo_b200/seal_format.py)."""
        global lw
        self._lw = lib if lib is not None else reinit_lw()
        self.option = option
        if path is None:
            path = str((Path.home() / ".hevm" / run_library.lower()).absolute())
        if not (Path(path) / "hevm_params.bin").is_file():
            # the reference blocks on input() here (runner.py:185-192); keys are cheap to
            # regenerate from a seed on the GPU, so we just create them.
            Path(path).mkdir(parents=True, exist_ok=True)
            self._lw.create_context(path.encode("utf-8"))
        enc = path.encode("utf-8")
        if option == "full":
            self.vm = self._lw.initFullVM(enc, run_hardware == "GPU")
        elif option == "client":
            self.vm = self._lw.initClientVM(enc)
        elif option == "server":
            self.vm = self._lw.initServerVM(enc)
        else:
            raise ValueError(option)
        self.slots = 1 << (self._lw.hevmx_param(self.vm, 0) - 1)
        self.hevm_path = ""
        if seal_dir is not None:
            from . import seal_format
            seal_format.load_seal_keys(self._lw, self.vm, seal_dir)

    def load(self, const_path, hevm_path, preprocess=True):
        if not Path(const_path).is_file():
            raise Exception(f"No file exists in const_path {const_path}")
        if not Path(hevm_path).is_file():
            raise Exception(f"No file exists in hevm_path {hevm_path}")
        self._lw.load(self.vm, str(const_path).encode("utf-8"), str(hevm_path).encode("utf-8"))
        if preprocess:
            self._lw.preprocess(self.vm)
        else:
            raise Exception("Not implemented in B200_HEVM")
        self.arglen = self._lw.getArgLen(self.vm)
        self.reslen = self._lw.getResLen(self.vm)
        self.hevm_path = str(hevm_path)

    def run(self):
        self._lw.run(self.vm)
        self._lw.printMem(self.vm)

    def setInput(self, i, data):
        if not isinstance(data, np.ndarray) or data.dtype != np.float64:
            data = np.array(data, dtype=np.float64)
        data = np.ascontiguousarray(data)
        carr = data.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        self._lw.encrypt(self.vm, i, carr, len(data))

    def setDebug(self, enable):
        self._lw.setDebug(self.vm, enable)

    def setToGPU(self, ongpu):
        self._lw.setToGPU(self.vm, ongpu)

    def getOutput(self):
        result = np.zeros((self.reslen, self.slots), dtype=np.float64)
        data = np.zeros(self.slots, dtype=np.float64)
        for i in range(self.reslen):
            carr = data.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            self._lw.decrypt_result(self.vm, i, carr)
            result[i] = data
        return result

    def printer(self, latency, rms, mem_usage=0.0):
        bench = re.search(r"optimized/(.*)/(.*)\.(.*)\._", self.hevm_path)
        print("======================================")
        print("---------------Option-----------------")
        if bench:
            print("compiler:", bench.group(1))
            print("benchname:", bench.group(2))
            print("waterline:", bench.group(3))
        print("library:", run_library)
        print("device:", run_hardware)
        print("---------------Result-----------------")
        print("latency:", latency)
        print("rms:", rms)
        print("======================================")
        print()
