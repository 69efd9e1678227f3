"""CPU tier: the CUDA warp-job bodies, replayed lane-by-lane by the test-only warp emulator
(tests/emul/warp_emul.cpp), must reproduce the oracle bit-for-bit at the real ring sizes
N = 2^14, 2^15 and 2^16.  This checks the kernels' index maths / fusion logic without a GPU; the `-m gpu`
tests then check the same kernels on the device through the C ABI.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from dacapo_b200 import hevm_asm as asm
from util import VM

HERE = Path(__file__).resolve().parent
u64p = C.POINTER(C.c_uint64)
NPR = 4


@pytest.fixture(scope="module", params=[15, 14, 16, 17])
def LOGN(request):
    return request.param


@pytest.fixture(scope="module")
def emul(LOGN):
    so = HERE / "emul" / "libwarp_emul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", str(so), str(HERE / "emul" / "warp_emul.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emul_create.restype = C.c_void_p
    lib.emul_create.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.emul_primes.argtypes = [C.c_void_p, u64p, u64p]
    lib.emul_ntt.argtypes = [C.c_void_p, u64p, C.c_int, C.c_int, C.c_int]
    lib.emul_keyswitch.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.c_uint32]
    lib.emul_rescale.argtypes = [C.c_void_p, u64p, u64p, C.c_int]
    lib.emul_keyswitch_sharded.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.c_uint32, C.c_int]
    lib.emul_keyswitch_sharded_split.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.c_uint32, C.c_int]
    lib.emul_set_fused.argtypes = [C.c_void_p, C.c_int]
    lib.emul_set_group_warps.argtypes = [C.c_void_p, C.c_int]
    lib.emul_waits_checked.argtypes = [C.c_void_p]
    lib.emul_waits_checked.restype = C.c_long
    lib.emul_keyswitch_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.POINTER(C.c_uint32)]
    return lib, lib.emul_create(LOGN, NPR, 60)


@pytest.fixture(scope="module")
def vm(oracle_lib, LOGN):
    return VM(oracle_lib, LOGN, NPR, nct=4, npt=1)


def _p(a):
    return a.ctypes.data_as(u64p)


def test_host_params_match_oracle(emul, vm):
    lib, h = emul
    q, psi = np.zeros(NPR, dtype=np.uint64), np.zeros(NPR, dtype=np.uint64)
    lib.emul_primes(h, _p(q), _p(psi))
    assert [int(x) for x in q] == vm.primes
    assert [int(x) for x in psi] == vm.roots


@pytest.mark.parametrize("prime", [0, NPR - 1])
def test_ntt_kernels_vs_oracle(emul, vm, prime):
    lib, h = emul
    rng = np.random.default_rng(prime)
    f = rng.integers(0, vm.primes[prime], size=(2, vm.N), dtype=np.uint64)
    f[1, :4] = [0, 1, vm.primes[prime] - 1, vm.primes[prime] - 2]
    exp = vm.ntt(f, prime)
    got = f.copy()
    lib.emul_ntt(h, _p(got), prime, 2, 0)
    assert np.array_equal(got, exp)
    lib.emul_ntt(h, _p(got), prime, 2, 1)
    assert np.array_equal(got, f)


def test_rescale_kernels_vs_oracle(emul, vm):
    lib, h = emul
    a = vm.random_ct(3, 1)
    vm.ct_write(0, a)
    vm.exec(asm.RESCALE, 1, 0)
    exp = vm.ct_read(1)
    got = np.zeros((2, 2, vm.N), dtype=np.uint64)
    lib.emul_rescale(h, _p(a), _p(got), 3)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("lvl", [1, 3])
def test_mulcc_relin_kernels_vs_oracle(emul, vm, lvl):
    lib, h = emul
    a, b = vm.random_ct(lvl, 2), vm.random_ct(lvl, 3)
    vm.ct_write(0, a)
    vm.ct_write(1, b)
    vm.exec(asm.MULCC, 2, 0, 1)
    exp = vm.ct_read(2)
    key = vm.key(2)
    got = np.zeros_like(a)
    lib.emul_keyswitch(h, 2, _p(a), _p(b), _p(got), lvl, _p(key), 0)
    assert np.array_equal(got, exp)
    # in place (dst aliases lhs), and squaring (lhs is rhs)
    a2 = a.copy()
    lib.emul_keyswitch(h, 2, _p(a2), _p(b), _p(a2), lvl, _p(key), 0)
    assert np.array_equal(a2, exp)
    vm.exec(asm.MULCC, 2, 0, 0)
    a3 = a.copy()
    lib.emul_keyswitch(h, 2, _p(a3), _p(a3), _p(a3), lvl, _p(key), 0)
    assert np.array_equal(a3, vm.ct_read(2))


@pytest.mark.parametrize("step", [1, -4])
def test_rotate_kernels_vs_oracle(emul, vm, step):
    lib, h = emul
    lvl = 2
    a = vm.random_ct(lvl, 4)
    vm.ct_write(0, a)
    vm.exec(asm.ROTATE, 1, 0, step)
    exp = vm.ct_read(1)
    elt = vm.lib.hevmx_galois_elt(vm.vm, step)
    key = vm.key(3, elt)
    got = a.copy()  # in place
    lib.emul_keyswitch(h, 1, _p(got), None, _p(got), lvl, _p(key), elt)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("group_warps", [1, 148, 400, 100000])
def test_target_groups_do_not_change_the_result(emul, vm, group_warps):
    """ops.hpp pick_groups: the fused inverse + forward pass-A launches are cut into 1 .. l+1 target groups depending on
    how many independent ops are in flight (vm.cu GroupWarpsScope); rotate, multiply + relinearise and rescale must give
    the oracle's words for every split."""
    lib, h = emul
    lvl = 3
    lib.emul_set_group_warps(h, group_warps)
    try:
        a, b = vm.random_ct(lvl, 31), vm.random_ct(lvl, 32)
        vm.ct_write(0, a)
        vm.ct_write(1, b)
        vm.exec(asm.ROTATE, 2, 0, 1)
        elt = vm.lib.hevmx_galois_elt(vm.vm, 1)
        got = np.zeros_like(a)
        lib.emul_keyswitch(h, 1, _p(a), None, _p(got), lvl, _p(vm.key(3, elt)), elt)
        assert np.array_equal(got, vm.ct_read(2))
        vm.exec(asm.MULCC, 2, 0, 1)
        got = np.zeros_like(a)
        lib.emul_keyswitch(h, 2, _p(a), _p(b), _p(got), lvl, _p(vm.key(2)), 0)
        assert np.array_equal(got, vm.ct_read(2))
        vm.exec(asm.RESCALE, 3, 2)
        out = np.zeros((2, lvl - 1, vm.N), dtype=np.uint64)
        lib.emul_rescale(h, _p(got), _p(out), lvl)
        assert np.array_equal(out, vm.ct_read(3))
    finally:
        lib.emul_set_group_warps(h, 1184)


@pytest.mark.parametrize("ranks", [1, 2, 3, 5])
def test_sharded_rotate_kernels_vs_oracle(emul, vm, ranks):
    """Limb-sharded key switch (targets partitioned over `ranks`, stages separated by the exchanges) == SEAL rotate."""
    lib, h = emul
    lvl, step = 3, 2
    a = vm.random_ct(lvl, 11)
    vm.ct_write(0, a)
    vm.exec(asm.ROTATE, 1, 0, step)
    exp = vm.ct_read(1)
    elt = vm.lib.hevmx_galois_elt(vm.vm, step)
    key = vm.key(3, elt)
    got = np.zeros_like(a)
    lib.emul_keyswitch_sharded(h, 1, _p(a), None, _p(got), lvl, _p(key), elt, ranks)
    assert np.array_equal(got, exp)
    got2 = np.zeros_like(a)  # stage 2 cut per source rank (the peer-to-peer path's order)
    lib.emul_keyswitch_sharded_split(h, 1, _p(a), None, _p(got2), lvl, _p(key), elt, ranks)
    assert np.array_equal(got2, exp)


@pytest.mark.parametrize("ranks", [2, 3])
def test_sharded_mulcc_kernels_vs_oracle(emul, vm, ranks):
    """Limb-sharded multiply + relinearise == SEAL multiply + relinearize_inplace, also in place."""
    lib, h = emul
    lvl = 3
    a, b = vm.random_ct(lvl, 21), vm.random_ct(lvl, 22)
    vm.ct_write(0, a)
    vm.ct_write(1, b)
    vm.exec(asm.MULCC, 2, 0, 1)
    exp = vm.ct_read(2)
    key = vm.key(2)
    got = np.zeros_like(a)
    lib.emul_keyswitch_sharded(h, 2, _p(a), _p(b), _p(got), lvl, _p(key), 0, ranks)
    assert np.array_equal(got, exp)
    a2 = a.copy()
    lib.emul_keyswitch_sharded(h, 2, _p(a2), _p(b), _p(a2), lvl, _p(key), 0, ranks)
    assert np.array_equal(a2, exp)


def test_fused_single_launch_vs_oracle(emul, vm):
    """ks_fused.cuh: the single-launch key switch / rescale (ticket-ordered units + completion counters).  The emulator
    replays the tickets in order and aborts if any unit's wait is not already satisfied by lower tickets; results must
    equal the oracle's rotate / multiply+relinearize / rescale, also in place and for a batch with different Galois keys."""
    lib, h = emul
    lib.emul_set_fused(h, 1)
    try:
        w0 = lib.emul_waits_checked(h)
        for lvl, step in ((3, 1), (1, -4), (2, 2)):
            a = vm.random_ct(lvl, 40 + lvl)
            vm.ct_write(0, a)
            vm.exec(asm.ROTATE, 1, 0, step)
            elt = vm.lib.hevmx_galois_elt(vm.vm, step)
            got = a.copy()
            lib.emul_keyswitch(h, 1, _p(got), None, _p(got), lvl, _p(vm.key(3, elt)), elt)
            assert np.array_equal(got, vm.ct_read(1)), (lvl, step)
        for lvl in (1, 3):
            a, b = vm.random_ct(lvl, 50 + lvl), vm.random_ct(lvl, 60 + lvl)
            vm.ct_write(0, a)
            vm.ct_write(1, b)
            vm.exec(asm.MULCC, 2, 0, 1)
            a2 = a.copy()
            lib.emul_keyswitch(h, 2, _p(a2), _p(b), _p(a2), lvl, _p(vm.key(2)), 0)
            assert np.array_equal(a2, vm.ct_read(2)), lvl
            vm.exec(asm.MULCC, 2, 0, 0)
            a3 = a.copy()
            lib.emul_keyswitch(h, 2, _p(a3), _p(a3), _p(a3), lvl, _p(vm.key(2)), 0)
            assert np.array_equal(a3, vm.ct_read(2)), lvl
        for lvl in (2, 3):
            a = vm.random_ct(lvl, 70 + lvl)
            vm.ct_write(0, a)
            vm.exec(asm.RESCALE, 1, 0)
            got = np.zeros((2, lvl - 1, vm.N), dtype=np.uint64)
            lib.emul_rescale(h, _p(a), _p(got), lvl)
            assert np.array_equal(got, vm.ct_read(1)), lvl
        assert lib.emul_waits_checked(h) > w0
        # batch of three ciphertexts with ONE key (same rotation step) in one launch
        lvl, step, n = 2, 1, 3
        elt = vm.lib.hevmx_galois_elt(vm.vm, step)
        cts = np.stack([vm.random_ct(lvl, 80 + k) for k in range(n)])
        exp = []
        for k in range(n):
            vm.ct_write(0, cts[k])
            vm.exec(asm.ROTATE, 1, 0, step)
            exp.append(vm.ct_read(1))
        got = np.zeros_like(cts)
        elts = (C.c_uint32 * n)(*([elt] * n))
        lib.emul_keyswitch_batch(h, 1, n, _p(cts), None, _p(got), lvl, _p(vm.key(3, elt)), elts)
        assert np.array_equal(got, np.stack(exp))
    finally:
        lib.emul_set_fused(h, 0)


def test_more_than_15_digits_vs_oracle(oracle_lib):
    """Levels above MAC_FLUSH_DIGITS = 15 take a second accumulation chunk in the key inner product (body_mac_dot):
    N = 2^14, 18 primes, level 17 -- rotate, multiply + relinearise (separate launches, single-launch form and the
    limb-sharded stages) against the oracle."""
    logn, npr, lvl = 14, 18, 17
    so = HERE / "emul" / "libwarp_emul.so"
    lib = C.CDLL(str(so))
    lib.emul_create.restype = C.c_void_p
    lib.emul_create.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.emul_keyswitch.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.c_uint32]
    lib.emul_keyswitch_sharded.argtypes = [C.c_void_p, C.c_int, u64p, u64p, u64p, C.c_int, u64p, C.c_uint32, C.c_int]
    lib.emul_set_fused.argtypes = [C.c_void_p, C.c_int]
    lib.emul_set_group_warps.argtypes = [C.c_void_p, C.c_int]
    h = lib.emul_create(logn, npr, 60)
    vm = VM(oracle_lib, logn, npr, nct=4, npt=1, galois_steps=(1,))
    a, b = vm.random_ct(lvl, 91), vm.random_ct(lvl, 92)
    vm.ct_write(0, a)
    vm.ct_write(1, b)
    vm.exec(asm.ROTATE, 2, 0, 1)
    vm.exec(asm.MULCC, 3, 0, 1)
    elt = vm.lib.hevmx_galois_elt(vm.vm, 1)
    gk, rk = vm.key(3, elt), vm.key(2)
    for fused in (0, 1):
        lib.emul_set_fused(h, fused)
        got = np.zeros_like(a)
        lib.emul_keyswitch(h, 1, _p(a), None, _p(got), lvl, _p(gk), elt)
        assert np.array_equal(got, vm.ct_read(2)), ("rotate", fused)
        got = np.zeros_like(a)
        lib.emul_keyswitch(h, 2, _p(a), _p(b), _p(got), lvl, _p(rk), 0)
        assert np.array_equal(got, vm.ct_read(3)), ("mulcc", fused)
    lib.emul_set_fused(h, 0)
    got = np.zeros_like(a)
    lib.emul_keyswitch_sharded(h, 1, _p(a), None, _p(got), lvl, _p(gk), elt, 4)
    assert np.array_equal(got, vm.ct_read(2)), "sharded rotate"


def test_shared_memory_layouts_are_bank_conflict_free(emul):
    """The padded layouts of ntt_core.cuh / ntt_bodies.cuh (TWB_ROW / twB_pos, MAC_XROW / xpad), checked with the hardware's
    rule for 16-byte shared-memory accesses: a quarter-warp (8 lanes) is served in one wavefront iff its 8 x 16 bytes
    fall into 32 distinct 4-byte banks.  Covers the accesses ncu attributed the round-1 conflicts to: the twiddle reads of
    the last two pass-B stages (layout C: lane holds indices 8 lane .. 8 lane + 7) and the stores / loads of the
    transformed digit rows in the key inner product."""
    lib, _ = emul
    lib.emul_xpad.argtypes = [C.c_int]
    lib.emul_twB_pos.argtypes = [C.c_int, C.c_int]
    lib.emul_layout_const.argtypes = [C.c_int]
    twb_row, xrow = lib.emul_layout_const(0), lib.emul_layout_const(1)

    def conflict_free(byte_addrs):  # 8 lanes x 16 bytes
        banks = [((a // 4) + j) % 32 for a in byte_addrs for j in range(4)]
        return len(set(banks)) == 32 and all(a % 16 == 0 for a in byte_addrs)

    # twiddle entries are 16 bytes; stage k = 6: lane reads g = 2 lane + j, stage 7: g = 4 lane + j
    for k, per_lane in ((6, 2), (7, 4)):
        used = set()
        for j in range(per_lane):
            for q in range(4):
                lanes = range(8 * q, 8 * q + 8)
                assert conflict_free([16 * lib.emul_twB_pos(k, per_lane * l + j) for l in lanes]), (k, j, q)
        for g in range(1 << k):
            used.add(lib.emul_twB_pos(k, g))
        assert len(used) == 1 << k and max(used) < twb_row
    # every (stage, group) has an entry of its own inside the row
    allpos = [lib.emul_twB_pos(k, g) for k in range(8) for g in range(1 << k)]
    assert len(set(allpos)) == 255 and max(allpos) < twb_row and min(allpos) == 0
    # digit rows: phase 1 stores words 8 lane + e (four 16-byte stores per lane), phase 2 loads words 2 tid, 2 tid + 1
    for e in (0, 2, 4, 6):
        for q in range(4):
            assert conflict_free([8 * (lib.emul_xpad(8 * l) + e) for l in range(8 * q, 8 * q + 8)]), ("store", e, q)
    for q in range(16):
        assert conflict_free([8 * lib.emul_xpad(2 * t) for t in range(8 * q, 8 * q + 8)]), ("load", q)
    assert all(lib.emul_xpad(i + 1) == lib.emul_xpad(i) + 1 for i in range(0, 256, 2))  # pairs stay adjacent (16-byte accesses)
    assert lib.emul_xpad(255) < xrow
