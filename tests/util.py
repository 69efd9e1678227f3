"""Shared helpers for the parity tests (test-only)."""
import ctypes as C
import os
import tempfile

import numpy as np

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def make_vm(lib, logn, nprimes, seed=0xDACA90, bits=60, keydir=None, galois_steps=None, env=None):
    """create_context + initFullVM on `lib` with the given ring geometry.  `galois_steps`: generate Galois keys for
    these rotation steps only (HEVM_GALOIS_STEPS, SEAL's create_galois_keys(steps)); default = SEAL's default set."""
    d = keydir or tempfile.mkdtemp(prefix="hevm_keys_")
    if not os.path.isfile(os.path.join(d, "hevm_params.bin")):
        old = {k: os.environ.get(k) for k in ("HEVM_LOGN", "HEVM_NUM_PRIMES", "HEVM_SEED", "HEVM_PRIME_BITS")}
        os.environ.update(HEVM_LOGN=str(logn), HEVM_NUM_PRIMES=str(nprimes), HEVM_SEED=str(seed), HEVM_PRIME_BITS=str(bits))
        try:
            lib.create_context(d.encode())
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    extra = dict(env or {})  # e.g. HEVM_SHARD_RANK / HEVM_SHARD_WORLD / HEVM_STREAMS, only while the VM is created
    if galois_steps is not None:
        extra["HEVM_GALOIS_STEPS"] = ",".join(str(int(x)) for x in galois_steps)
    old = {k: os.environ.get(k) for k in extra}
    os.environ.update({k: str(v) for k, v in extra.items()})
    try:
        vm = lib.initFullVM(d.encode(), True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return vm, d


class VM:
    """Thin convenience wrapper over the hevmx_* hooks of one library."""

    def __init__(self, lib, logn, nprimes, seed=0xDACA90, keydir=None, nct=8, npt=4, galois_steps=None, env=None):
        self.lib = lib
        self.vm, self.keydir = make_vm(lib, logn, nprimes, seed, keydir=keydir, galois_steps=galois_steps, env=env)
        self.logn, self.N, self.L = logn, 1 << logn, nprimes
        p = np.zeros(nprimes, dtype=np.uint64)
        lib.hevmx_primes(self.vm, p.ctypes.data_as(_u64p))
        self.primes = [int(x) for x in p]
        r = np.zeros(nprimes, dtype=np.uint64)
        lib.hevmx_roots(self.vm, r.ctypes.data_as(_u64p))
        self.roots = [int(x) for x in r]
        lib.hevmx_resize(self.vm, nct, npt)

    def ct_info(self, reg):
        lv, sc = C.c_int64(), C.c_double()
        self.lib.hevmx_ct_info(self.vm, reg, C.byref(lv), C.byref(sc))
        return lv.value, sc.value

    def ct_read(self, reg):
        lv, _ = self.ct_info(reg)
        out = np.zeros((2, lv, self.N), dtype=np.uint64)
        self.lib.hevmx_ct_read(self.vm, reg, out.ctypes.data_as(_u64p))
        return out

    def ct_write(self, reg, arr, scale=2.0 ** 40):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        assert arr.shape[0] == 2 and arr.shape[2] == self.N
        self.lib.hevmx_ct_write(self.vm, reg, arr.ctypes.data_as(_u64p), arr.shape[1], scale)

    def pt_info(self, reg):
        lv, sc = C.c_int64(), C.c_double()
        self.lib.hevmx_pt_info(self.vm, reg, C.byref(lv), C.byref(sc))
        return lv.value, sc.value

    def pt_read(self, reg):
        lv, _ = self.pt_info(reg)
        out = np.zeros((lv, self.N), dtype=np.uint64)
        self.lib.hevmx_pt_read(self.vm, reg, out.ctypes.data_as(_u64p))
        return out

    def pt_write(self, reg, arr, scale=2.0 ** 40):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        self.lib.hevmx_pt_write(self.vm, reg, arr.ctypes.data_as(_u64p), arr.shape[0], scale)

    def exec(self, opcode, dst, lhs=0, rhs=0, sync=True):
        self.lib.hevmx_exec(self.vm, opcode, dst, lhs, rhs & 0xFFFF)
        if sync:
            self.lib.hevmx_sync(self.vm)

    def exec_batch(self, opcode, dst, lhs, rhs, sync=True):
        """n independent ops of one opcode in one batched launch (libB200_HEVM.so only)."""
        arr = [np.ascontiguousarray(np.asarray(v, dtype=np.int64) & (0xFFFF if k == 2 else -1)) for k, v in enumerate((dst, lhs, rhs))]
        i64p = C.POINTER(C.c_int64)
        self.lib.hevmx_exec_batch(self.vm, opcode, len(arr[0]), *(a.ctypes.data_as(i64p) for a in arr))
        if sync:
            self.lib.hevmx_sync(self.vm)

    def ntt(self, data, prime_idx, inverse=False):
        a = np.ascontiguousarray(data, dtype=np.uint64).copy()
        self.lib.hevmx_ntt(self.vm, a.ctypes.data_as(_u64p), prime_idx, a.size // self.N, int(inverse))
        return a

    def encode(self, ptreg, vals, level, scale_bits):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self.lib.hevmx_encode(self.vm, ptreg, v.ctypes.data_as(_f64p), v.size, level, scale_bits)

    def decode(self, ptreg):
        out = np.zeros(self.N // 2, dtype=np.float64)
        self.lib.hevmx_decode(self.vm, ptreg, out.ctypes.data_as(_f64p))
        return out

    def encrypt_pt(self, ptreg, ctreg, counter=None):
        if counter is not None:
            self.lib.hevmx_set_enc_counter(self.vm, counter)
        self.lib.hevmx_encrypt_pt(self.vm, ptreg, ctreg)

    def decrypt_to_pt(self, ctreg, ptreg):
        self.lib.hevmx_decrypt_to_pt(self.vm, ctreg, ptreg)

    def decrypt_decode(self, ctreg, ptreg=0):
        self.decrypt_to_pt(ctreg, ptreg)
        return self.decode(ptreg)

    def key(self, which, elt=0):
        n = self.lib.hevmx_key_read(self.vm, which, elt, None)
        if n < 0:
            return None
        out = np.zeros(n, dtype=np.uint64)
        self.lib.hevmx_key_read(self.vm, which, elt, out.ctypes.data_as(_u64p))
        return out

    def random_ct(self, level, seed):
        rng = np.random.default_rng(seed)
        a = np.zeros((2, level, self.N), dtype=np.uint64)
        for i in range(level):
            a[:, i, :] = rng.integers(0, self.primes[i], size=(2, self.N), dtype=np.uint64)
        return a

    def random_pt(self, level, seed):
        rng = np.random.default_rng(seed)
        a = np.zeros((level, self.N), dtype=np.uint64)
        for i in range(level):
            a[i, :] = rng.integers(0, self.primes[i], size=self.N, dtype=np.uint64)
        return a
