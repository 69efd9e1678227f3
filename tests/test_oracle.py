"""CPU tests: pin the oracle (oracle/) against independent Python big-integer maths.

The reference holds no golden vectors for this path ("parity
unpinned" at the SEAL boundary), so the pins are: externally checkable constants,
definition-level Python restatements on small rings, and algebraic identities.
"""
import numpy as np
import pytest

import pyref
from dacapo_b200 import hevm_asm as asm
from util import VM

LOGN, NPR = 8, 4  # small ring for O(N^2) definition checks


@pytest.fixture(scope="module")
def vm(oracle_lib):
    return VM(oracle_lib, LOGN, NPR, nct=8, npt=4)


# ---- parameters -------------------------------------------------------------------------------
def test_seal_prime_chain_n15():
    """14 x 60-bit primes for N=2^15; q_13 = 0xFFFFFFFFFFC0001."""
    ps = pyref.seal_primes(1 << 15, 60, 14)
    deltas = [(1 << 60) - p for p in ps]
    assert deltas == [25427967, 24969215, 23199743, 21299199, 20316159, 16908287, 14417919, 14024703,
                      13434879, 11206655, 10878975, 9830399, 8126463, 262143]
    assert ps[13] == 0xFFFFFFFFFFC0001
    assert all(p % (1 << 16) == 1 for p in ps)


def test_oracle_primes_and_roots(vm):
    assert vm.primes == pyref.seal_primes(vm.N, 60, NPR)
    for q, psi in zip(vm.primes, vm.roots):
        assert psi == pyref.minimal_root(vm.N, q)
        assert pow(psi, vm.N, q) == q - 1


def test_oracle_primes_n15_golden(oracle_lib):
    """Full-size geometry: primes and minimal roots of the reference configuration (golden file)."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "params_n15.json").read_text())
    assert g["primes"] == pyref.seal_primes(1 << 15, 60, 14)
    # roots were produced by pyref.minimal_root (tests/golden/make_golden.py); spot check two here
    for i in (0, 13):
        assert g["roots"][i] == pyref.minimal_root(1 << 15, g["primes"][i])


# ---- NTT ------------------------------------------------------------------------------------------
def test_ntt_matches_definition(vm):
    rng = np.random.default_rng(1)
    for i, (q, psi) in enumerate(zip(vm.primes, vm.roots)):
        f = rng.integers(0, q, size=vm.N, dtype=np.uint64)
        got = vm.ntt(f, i)
        exp = pyref.ntt_eval([int(x) for x in f], psi, q)
        assert [int(x) for x in got] == exp
        back = vm.ntt(got, i, inverse=True)
        assert np.array_equal(back, f)


def test_ntt_edge_values(vm):
    q, psi = vm.primes[0], vm.roots[0]
    for f in ([0] * vm.N, [q - 1] * vm.N, [1] + [0] * (vm.N - 1), [0] * (vm.N - 1) + [q - 1]):
        got = vm.ntt(np.array(f, dtype=np.uint64), 0)
        assert [int(x) for x in got] == pyref.ntt_fast(f, psi, q)


def test_ntt_n15_fast(oracle_lib):
    big = VM(oracle_lib, 12, 3, nct=2, npt=1)
    rng = np.random.default_rng(2)
    q, psi = big.primes[2], big.roots[2]
    f = rng.integers(0, q, size=big.N, dtype=np.uint64)
    assert [int(x) for x in big.ntt(f, 2)] == pyref.ntt_fast([int(x) for x in f], psi, q)


# ---- evaluator ops vs python maths ----------------------------------------------------------------
def test_elementwise_ops(vm):
    lvl = 3
    a, b = vm.random_ct(lvl, 10), vm.random_ct(lvl, 11)
    p = vm.random_pt(lvl, 12)
    vm.ct_write(0, a)
    vm.ct_write(1, b)
    vm.pt_write(0, p)
    qs = np.array(vm.primes[:lvl], dtype=object).reshape(1, lvl, 1)
    ao, bo, po = a.astype(object), b.astype(object), p.astype(object)[None]
    vm.exec(asm.ADDCC, 2, 0, 1)
    assert np.array_equal(vm.ct_read(2).astype(object), (ao + bo) % qs)
    vm.exec(asm.NEGATE, 2, 0)
    assert np.array_equal(vm.ct_read(2).astype(object), (-ao) % qs)
    vm.exec(asm.MULCP, 2, 0, 0)
    assert np.array_equal(vm.ct_read(2).astype(object), (ao * po) % qs)
    vm.exec(asm.ADDCP, 2, 0, 0)
    exp = ao.copy()
    exp[0] = (exp[0] + po[0]) % qs[0]
    assert np.array_equal(vm.ct_read(2).astype(object), exp)
    vm.exec(asm.MODSWITCH, 2, 0, 2)
    assert np.array_equal(vm.ct_read(2), a[:, :1, :])
    assert vm.ct_info(2)[0] == 1


def test_rescale_is_exact_rounding(vm):
    lvl = 3
    a = vm.random_ct(lvl, 20)
    vm.ct_write(0, a, scale=2.0 ** 80)
    vm.exec(asm.RESCALE, 1, 0)
    got = vm.ct_read(1)
    for j in range(2):
        exp = pyref.rescale_poly([[int(x) for x in a[j, i]] for i in range(lvl)], vm.primes[:lvl], vm.roots[:lvl])
        for i in range(lvl - 1):
            assert [int(x) for x in got[j, i]] == exp[i]
    lv, sc = vm.ct_info(1)
    assert lv == 2 and sc == 2.0 ** 80 / float(vm.primes[2])


def _key_as_lists(vm, words, L, N):
    k = words.reshape(L - 1, 2, L, N)
    return [[[[int(x) for x in k[J, K, I]] for I in range(L)] for K in range(2)] for J in range(L - 1)]


def test_mulcc_relin_matches_math(vm):
    lvl = 2
    a, b = vm.random_ct(lvl, 30), vm.random_ct(lvl, 31)
    vm.ct_write(0, a)
    vm.ct_write(1, b)
    vm.exec(asm.MULCC, 2, 0, 1)
    got = vm.ct_read(2)
    relin = _key_as_lists(vm, vm.key(2), vm.L, vm.N)
    qs = vm.primes[:lvl]
    d0 = [[int(a[0, i, k]) * int(b[0, i, k]) % qs[i] for k in range(vm.N)] for i in range(lvl)]
    d1 = [[(int(a[0, i, k]) * int(b[1, i, k]) + int(a[1, i, k]) * int(b[0, i, k])) % qs[i] for k in range(vm.N)] for i in range(lvl)]
    d2 = [[int(a[1, i, k]) * int(b[1, i, k]) % qs[i] for k in range(vm.N)] for i in range(lvl)]
    delta = pyref.switch_key(d2, relin, vm.primes, vm.roots, lvl)
    for i in range(lvl):
        assert [int(x) for x in got[0, i]] == [(d0[i][k] + delta[0][i][k]) % qs[i] for k in range(vm.N)]
        assert [int(x) for x in got[1, i]] == [(d1[i][k] + delta[1][i][k]) % qs[i] for k in range(vm.N)]


@pytest.mark.parametrize("step", [1, -1, 4, 3, -7])
def test_rotate_matches_math(vm, step):
    lvl = 2
    a = vm.random_ct(lvl, 40 + step)
    vm.ct_write(0, a)
    vm.exec(asm.ROTATE, 1, 0, step)
    got = vm.ct_read(1)
    qs = vm.primes[:lvl]
    cur = [[[int(x) for x in a[j, i]] for i in range(lvl)] for j in range(2)]
    elt_direct = pyref.galois_elt(step, vm.N)
    has_direct = vm.key(3, elt_direct) is not None
    terms = [step] if has_direct else [t for t in pyref.naf(step) if abs(t) != vm.N // 2]
    for t in terms:
        elt = pyref.galois_elt(t, vm.N)
        assert vm.lib.hevmx_galois_elt(vm.vm, t) == elt
        tab = pyref.galois_table(elt, vm.N)
        key = _key_as_lists(vm, vm.key(3, elt), vm.L, vm.N)
        p0 = [[cur[0][i][tab[k]] for k in range(vm.N)] for i in range(lvl)]
        p1 = [[cur[1][i][tab[k]] for k in range(vm.N)] for i in range(lvl)]
        delta = pyref.switch_key(p1, key, vm.primes, vm.roots, lvl)
        cur = [[[(p0[i][k] + delta[0][i][k]) % qs[i] for k in range(vm.N)] for i in range(lvl)],
               [delta[1][i] for i in range(lvl)]]
    for j in range(2):
        for i in range(lvl):
            assert [int(x) for x in got[j, i]] == cur[j][i]


def test_galois_table_is_automorphism(vm):
    """apply_galois_ntt table == NTT of f(x^elt) (SURVEY A.2.11 ii) via definition-level maths."""
    q, psi, N = vm.primes[0], vm.roots[0], vm.N
    rng = np.random.default_rng(5)
    f = [int(x) for x in rng.integers(0, q, size=N, dtype=np.uint64)]
    F = pyref.ntt_fast(f, psi, q)
    for elt in (3, 2 * N - 1, pow(3, N // 2 - 1, 2 * N)):
        g = [0] * N
        for k, c in enumerate(f):
            e = k * elt % (2 * N)
            if e < N:
                g[e] = (g[e] + c) % q
            else:
                g[e - N] = (g[e - N] - c) % q
        tab = pyref.galois_table(elt, N)
        assert pyref.ntt_fast(g, psi, q) == [F[tab[i]] for i in range(N)]


# ---- keys ------------------------------------------------------------------------------------------
def test_key_structure(vm):
    """pk and key-switch keys decrypt to what SEAL's keygen defines."""
    L, N = vm.L, vm.N
    sk = vm.key(0).reshape(L, N)
    pk = vm.key(1).reshape(2, L, N)
    relin = vm.key(2).reshape(L - 1, 2, L, N)
    p = vm.primes[L - 1]
    for i in range(L):
        q, psi = vm.primes[i], vm.roots[i]
        s = [int(x) for x in sk[i]]
        coeff = pyref.intt_fast(s, psi, q)
        assert set(coeff) <= {0, 1, q - 1}
        e = pyref.intt_fast([(int(pk[0, i, k]) + int(pk[1, i, k]) * s[k]) % q for k in range(N)], psi, q)
        assert all(min(x, q - x) <= 21 for x in e)  # -(e): centred binomial, |e| <= 21
    for J in range(L - 1):
        for i in range(L):
            q, psi = vm.primes[i], vm.roots[i]
            s = [int(x) for x in sk[i]]
            v = [(int(relin[J, 0, i, k]) + int(relin[J, 1, i, k]) * s[k]) % q for k in range(N)]
            if i == J:
                v = [(v[k] - (p % q) * s[k] * s[k]) % q for k in range(N)]
            e = pyref.intt_fast(v, psi, q)
            assert all(min(x, q - x) <= 21 for x in e)
    assert vm.lib.hevmx_param(vm.vm, 5) == 2 * (vm.logn - 1)  # default set: 3^(+-2^i) (the last pair coincides) + conjugation


# ---- encoder / encryptor semantics ----------------------------------------------------------------
def test_encode_decode_roundtrip(vm):
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, vm.N // 2)
    vm.encode(0, x, 3, 40)
    assert vm.pt_info(0) == (3, 2.0 ** 40)
    y = vm.decode(0)
    assert np.max(np.abs(x - y)) < 1e-9


def test_encode_slot_semantics(vm):
    """Plaintext polynomial evaluated at zeta^(3^j) equals scale*x_j (canonical embedding)."""
    x = np.zeros(vm.N // 2)
    x[1] = 1.0
    vm.encode(0, x, 1, 30)
    q, psi = vm.primes[0], vm.roots[0]
    coeff = pyref.intt_fast([int(v) for v in vm.pt_read(0)[0]], psi, q)
    centred = np.array([c if c < q // 2 else c - q for c in coeff], dtype=np.float64)
    zeta = np.exp(1j * np.pi / vm.N)
    for j in (0, 1, 2, 5):
        z = zeta ** pow(3, j, 2 * vm.N)
        val = np.sum(centred * z ** np.arange(vm.N)) / 2.0 ** 30
        assert abs(val - x[j]) < 1e-6


def test_encrypt_decrypt_and_homomorphic_identities(vm):
    rng = np.random.default_rng(8)
    n = vm.N // 2
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    vm.encode(0, x, 3, 50)
    vm.encode(1, y, 3, 50)
    vm.encrypt_pt(0, 0)
    vm.encrypt_pt(1, 1)
    assert np.max(np.abs(vm.decrypt_decode(0, 2) - x)) < 1e-7
    vm.exec(asm.ADDCC, 2, 0, 1)
    assert np.max(np.abs(vm.decrypt_decode(2, 2) - (x + y))) < 1e-7
    vm.exec(asm.MULCC, 3, 0, 1)
    vm.exec(asm.RESCALE, 3, 3)
    assert vm.ct_info(3)[0] == 2
    assert np.max(np.abs(vm.decrypt_decode(3, 2) - x * y)) < 1e-6
    vm.exec(asm.MULCP, 4, 0, 1)
    assert np.max(np.abs(vm.decrypt_decode(4, 2) - x * y)) < 1e-6
    for step in (1, -3, 5):
        vm.exec(asm.ROTATE, 5, 0, step)
        assert np.max(np.abs(vm.decrypt_decode(5, 2) - np.roll(x, -step))) < 1e-6
    vm.exec(asm.BOOTSTRAP, 6, 3, 3)  # decrypt + re-encrypt at level 3
    lv, sc = vm.ct_info(6)
    assert lv == 3 and sc == 2.0 ** int(np.log2(vm.ct_info(3)[1]))
    assert np.max(np.abs(vm.decrypt_decode(6, 2) - x * y)) < 1e-5


def test_aliasing(vm):
    a, b = vm.random_ct(2, 50), vm.random_ct(2, 51)
    for op, args in ((asm.ADDCC, (0, 1)), (asm.MULCC, (0, 1)), (asm.MULCC, (0, 0)), (asm.ROTATE, (0, 3)), (asm.RESCALE, (0, 0))):
        vm.ct_write(0, a)
        vm.ct_write(1, b)
        vm.exec(op, 2, *args)
        exp = vm.ct_read(2)
        for dst in (0, 1):
            if op in (asm.ROTATE, asm.RESCALE) and dst == 1:
                continue
            vm.ct_write(0, a)
            vm.ct_write(1, b)
            vm.exec(op, dst, *args)
            assert np.array_equal(vm.ct_read(dst), exp)


# ---- program container ------------------------------------------------------------------------------
def test_hevm_program_roundtrip(vm, tmp_path):
    p = asm.Program(init_level=3)
    x = p.arg(50, 3)
    t, u = p.new_ct(), p.new_ct()
    c = p.const([0.5, -0.25])
    pt = p.new_pt()
    ones = p.new_pt()
    p.encode(pt, c, 3, 50)
    p.encode(ones, -1, 2, 20)
    p.emit(asm.PLACEHOLDER, 0xBEEF, 0xDEAD, 0xF00D)
    p.emit(asm.MULCP, t, x, pt)
    p.emit(asm.RESCALE, t, t)
    p.rotate(u, t, -2)
    p.emit(asm.ADDCC, u, u, t)
    p.emit(asm.MULCP, u, u, ones)
    p.result(u, 40, 2)
    cst, hv = tmp_path / "a.cst", tmp_path / "a.hevm"
    p.save(cst, hv)
    q = asm.parse_hevm(hv.read_bytes())
    assert q.ops == p.ops and q.res_dst == p.res_dst and q.num_ct == p.num_ct
    assert [list(c) for c in asm.parse_cst(cst.read_bytes())] == [[0.5, -0.25]]
    lib = vm.lib
    lib.load(vm.vm, str(cst).encode(), str(hv).encode())
    lib.preprocess(vm.vm)
    assert lib.getArgLen(vm.vm) == 1 and lib.getResLen(vm.vm) == 1 and lib.getResIdx(vm.vm, 0) == u
    n = vm.N // 2
    data = np.linspace(-1, 1, n)
    import ctypes as C
    lib.encrypt(vm.vm, 0, data.ctypes.data_as(C.POINTER(C.c_double)), n)
    lib.run(vm.vm)
    out = np.zeros(n)
    lib.decrypt_result(vm.vm, 0, out.ctypes.data_as(C.POINTER(C.c_double)))
    w = data * np.tile([0.5, -0.25], n // 2)
    exp = np.roll(w, 2) + w
    assert np.max(np.abs(out - exp)) < 1e-5
    vm.lib.hevmx_resize(vm.vm, 8, 4)


# ---- known-answer tests for the externally documented SEAL facts the oracle relies on ----
def test_kat_galois_elt_from_step(vm):
    """SEAL GaloisTool::get_elt_from_step (util/galois.cpp): generator 3 modulo 2N; a positive step s (LEFT rotation) maps
    to 3^s, a negative one to 3^(N/2 - |s|), step 0 to the conjugation element 2N - 1.  N = 2^11 here: 2N = 4096."""
    m = 2 * vm.N
    g = lambda st: vm.lib.hevmx_galois_elt(vm.vm, st)
    assert g(1) == 3 and g(2) == 9 and g(3) == 27 and g(0) == m - 1
    assert g(-1) == pow(3, vm.N // 2 - 1, m) and (g(-1) * 3) % m == 1  # 3^(N/2) = 1 mod 2N: the inverse rotation
    assert g(5) == pow(3, 5, m) and g(-7) == pow(3, vm.N // 2 - 7, m)


def test_kat_constant_vector_encodes_to_constant_polynomial(vm):
    """CKKSEncoder: the all-c slot vector is the constant polynomial round(c * scale); in NTT form EVERY residue of limb i
    is that constant mod q_i.  Hand-computable: c = 0.75, scale 2^40 -> 0x0C000000000."""
    n = vm.N // 2
    vm.encode(0, np.full(n, 0.75), 3, 40)
    p = vm.pt_read(0)
    assert p.shape == (3, vm.N)
    for i in range(3):
        assert np.all(p[i] == np.uint64((3 << 38) % vm.primes[i]))
    vm.encode(0, np.full(n, -1.0), 2, 30)  # a negative constant is stored as q - |x|
    p = vm.pt_read(0)
    for i in range(2):
        assert np.all(p[i] == np.uint64(vm.primes[i] - (1 << 30)))


def test_kat_slot_zero_and_rotation_direction(vm):
    """SEAL Evaluator::rotate_vector(ct, k): k > 0 rotates the slot vector to the LEFT (slot i receives slot i + k).  A unit
    vector in slot 5 therefore shows up in slot 4 after rotate(1) and in slot 7 after rotate(-2); slot 0 wraps to slot n-1."""
    n = vm.N // 2
    e5, e0 = np.zeros(n), np.zeros(n)
    e5[5], e0[0] = 1.0, 1.0
    for vec, step, where in ((e5, 1, 4), (e5, -2, 7), (e0, 1, n - 1), (e0, -1, 1)):
        vm.encode(0, vec, 3, 40)
        vm.encrypt_pt(0, 0)
        vm.exec(asm.ROTATE, 1, 0, step)
        out = vm.decrypt_decode(1, 1)
        assert int(np.argmax(out)) == where and abs(out[where] - 1.0) < 1e-6 and np.sum(np.abs(out)) - abs(out[where]) < 1e-3
