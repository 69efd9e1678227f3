"""Diagnose test_seal_key_directory_round_trip: which op / which key differs between the exporting and the importing VM."""
import os, sys, tempfile
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from dacapo_b200 import _binding, hevm_asm as asm, seal_format as sf
from util import VM
lib = _binding.bind(_binding.B200_LIB)
steps = (1, -2, 64)
for trial in range(3):
    a = VM(lib, 14, 5, seed=0xA11CE, keydir=tempfile.mkdtemp(), nct=4, npt=2, galois_steps=steps)
    b = VM(lib, 14, 5, seed=0xB0B, keydir=tempfile.mkdtemp(), nct=4, npt=2, galois_steps=(1,))
    d = tempfile.mkdtemp()
    sf.export_vm_keys(lib, a.vm, d, sf.COMPR_ZSTD)
    sf.load_seal_keys(lib, b.vm, d)
    for which in (0, 1, 2):
        print(trial, "key", which, np.array_equal(a.key(which), b.key(which)))
    for st in steps:
        e = lib.hevmx_galois_elt(a.vm, st)
        print(trial, "galois", st, e, lib.hevmx_galois_elt(b.vm, st), np.array_equal(a.key(3, e), b.key(3, e)))
    x = np.random.default_rng(5).uniform(-1, 1, a.N // 2)
    a.encode(0, x, 4, 50)
    a.encrypt_pt(0, 0, counter=7)
    ct = a.ct_read(0)
    b.ct_write(0, ct, 2.0 ** 50)
    print(trial, "ct0", np.array_equal(a.ct_read(0), b.ct_read(0)), a.ct_info(0), b.ct_info(0))
    seq = [(asm.ROTATE, 1, 0, 64), (asm.ROTATE, 1, 1, -2), (asm.MULCC, 2, 1, 0), (asm.RESCALE, 2, 2, 0)]
    for op in seq:
        for vm in (a, b):
            vm.exec(*op)
        print(trial, "after", op, np.array_equal(a.ct_read(op[1]), b.ct_read(op[1])), a.ct_info(op[1]), b.ct_info(op[1]))
    # the same ops again on fresh registers, a second time
    for vm in (a, b):
        vm.exec(asm.ROTATE, 3, 0, 64)
    print(trial, "rot64 again", np.array_equal(a.ct_read(3), b.ct_read(3)))
    for vm in (a, b):
        vm.exec(asm.ROTATE, 3, 0, -2)
    print(trial, "rot-2 fresh", np.array_equal(a.ct_read(3), b.ct_read(3)))
    for vm in (a, b):
        vm.exec(asm.ROTATE, 3, 0, 1)
    print(trial, "rot1 fresh", np.array_equal(a.ct_read(3), b.ct_read(3)))
