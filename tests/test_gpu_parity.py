"""GPU tier (-m gpu): libB200_HEVM.so vs the CPU oracle, bit-exact, through the C ABI.

Full reference geometry: N = 2^15, 14 x 60-bit primes (SEAL_HEVM.cpp:39-53).  Both libraries
are initialised from the same parameter/seed file, so keys, encryptions and every opcode's
output residues must be identical words.  Floating-point (encode/decode) parity is also
required bit-exact; the tolerance for end-to-end decrypted values against plain numpy maths
is 1e-5 (the reference reports RMS ~1e-3 on ResNet, README.md:187).
"""
import ctypes as C

import numpy as np
import pytest

from dacapo_b200 import hevm_asm as asm
from util import VM

pytestmark = pytest.mark.gpu
LOGN, NPR = 15, 14


@pytest.fixture(scope="module")
def pair(oracle_lib, b200_lib, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("keys"))
    g = VM(b200_lib, LOGN, NPR, keydir=d, nct=8, npt=4)
    o = VM(oracle_lib, LOGN, NPR, keydir=d, nct=8, npt=4)
    assert b200_lib.hevmx_backend() == b"b200-cuda-sm_100a"
    return g, o


def both(pair, fn):
    g, o = pair
    return fn(g), fn(o)


def test_parameters(pair):
    g, o = pair
    assert g.primes == o.primes and g.roots == o.roots
    assert g.primes[13] == 0xFFFFFFFFFFC0001


def test_keys_bit_exact(pair):
    g, o = pair
    for which in (0, 1, 2):
        assert np.array_equal(g.key(which), o.key(which)), f"key kind {which}"
    for step in (1, -1, 64, -8192):
        elt = o.lib.hevmx_galois_elt(o.vm, step)
        assert g.lib.hevmx_galois_elt(g.vm, step) == elt
        assert np.array_equal(g.key(3, elt), o.key(3, elt)), f"galois key step {step}"
    assert g.lib.hevmx_param(g.vm, 5) == o.lib.hevmx_param(o.vm, 5) == 28


@pytest.mark.parametrize("prime", list(range(NPR)))  # every prime of the chain
def test_ntt_bit_exact(pair, prime):
    g, o = pair
    rng = np.random.default_rng(prime)
    q = o.primes[prime]
    f = rng.integers(0, q, size=(3, o.N), dtype=np.uint64)
    f[0, :6] = [0, 1, q - 1, q - 2, 2, q // 2]
    f[2, :] = q - 1
    F = o.ntt(f, prime)
    assert np.array_equal(g.ntt(f, prime), F)
    assert np.array_equal(g.ntt(F, prime, inverse=True), f)


@pytest.mark.parametrize("prime", [0, 6, 13])
def test_cluster_ntt_bit_exact(pair, prime):
    """The single-pass cluster NTT (ntt_cluster.cuh: TMA tile loads, DSMEM transpose inside an 8-CTA cluster) must give
    the same canonical residues as the oracle's forward NTT; 19 limbs = more clusters than fit the GPU at once."""
    g, o = pair
    rng = np.random.default_rng(50 + prime)
    q = o.primes[prime]
    f = rng.integers(0, q, size=(19, o.N), dtype=np.uint64)
    f[0, :4] = [0, 1, q - 1, q // 2]
    f[18, :] = q - 1
    assert np.array_equal(g.ntt(f, prime, inverse=2), o.ntt(f, prime))


@pytest.mark.parametrize("lvl", list(range(1, NPR)))  # every level of the modulus chain
def test_elementwise_ops(pair, lvl):
    g, o = pair
    a, b, p = o.random_ct(lvl, lvl), o.random_ct(lvl, 100 + lvl), o.random_pt(lvl, 200 + lvl)
    for vm in pair:
        vm.ct_write(0, a, 2.0 ** 40)
        vm.ct_write(1, b, 2.0 ** 41)
        vm.pt_write(0, p, 2.0 ** 30)
    for op, args in ((asm.ADDCC, (0, 1)), (asm.NEGATE, (0, 0)), (asm.ADDCP, (0, 0)), (asm.MULCP, (0, 0))):
        for vm in pair:
            vm.exec(op, 2, *args)
        assert np.array_equal(g.ct_read(2), o.ct_read(2)), asm.OP_NAMES[op]
        assert g.ct_info(2) == o.ct_info(2)
        assert g.ct_info(0) == o.ct_info(0)  # addcc/addcp overwrite the lhs scale (SEAL_HEVM.cpp:301,308)
    if lvl > 2:
        for vm in pair:
            vm.exec(asm.MODSWITCH, 3, 0, 2)
            vm.exec(asm.MODSWITCH, 0, 0, 1)  # in place
        assert np.array_equal(g.ct_read(3), o.ct_read(3)) and g.ct_info(3) == o.ct_info(3) == (lvl - 2, 2.0 ** 30)
        assert np.array_equal(g.ct_read(0), o.ct_read(0)) and g.ct_info(0)[0] == lvl - 1


@pytest.mark.parametrize("lvl", list(range(2, NPR)))
def test_rescale(pair, lvl):
    g, o = pair
    a = o.random_ct(lvl, 300 + lvl)
    for dst in (1, 0):  # out of place, in place
        for vm in pair:
            vm.ct_write(0, a, 2.0 ** 100)
            vm.exec(asm.RESCALE, dst, 0)
        assert np.array_equal(g.ct_read(dst), o.ct_read(dst))
        assert g.ct_info(dst) == o.ct_info(dst)


@pytest.mark.parametrize("lvl", list(range(1, NPR)))
def test_mulcc_relin(pair, lvl):
    g, o = pair
    a, b = o.random_ct(lvl, 400 + lvl), o.random_ct(lvl, 500 + lvl)
    forms = ((2, 0, 1), (0, 0, 1), (1, 0, 1), (2, 0, 0), (0, 0, 0)) if lvl in (1, 2, 6, 13) else ((2, 0, 1), (0, 0, 0))
    for dst, lhs, rhs in forms:
        for vm in pair:
            vm.ct_write(0, a, 2.0 ** 40)
            vm.ct_write(1, b, 2.0 ** 40)
            vm.exec(asm.MULCC, dst, lhs, rhs)
        assert np.array_equal(g.ct_read(dst), o.ct_read(dst)), (dst, lhs, rhs)
        assert g.ct_info(dst) == o.ct_info(dst)


@pytest.mark.parametrize("lvl,step", [(1, 1), (2, -1), (5, 128), (13, 1), (13, -8192), (4, 12285), (3, -15360), (2, 0), (13, 16383)]
                         + [(l, (1, -2, 4096)[l % 3]) for l in range(3, 13)])  # + every remaining level with a single-key step
def test_rotate(pair, lvl, step):
    g, o = pair
    a = o.random_ct(lvl, 600 + lvl)
    for dst in (1, 0):
        for vm in pair:
            vm.ct_write(0, a, 2.0 ** 40)
            vm.exec(asm.ROTATE, dst, 0, step)
        assert np.array_equal(g.ct_read(dst), o.ct_read(dst)), (lvl, step, dst)


@pytest.mark.parametrize("lvl,step,ranks", [(13, 1, 1), (13, -8192, 2), (7, 64, 3), (2, 1, 4), (1, 1, 2)])
def test_sharded_keyswitch_stages_bit_exact(pair, lvl, step, ranks):
    """Limb-sharded rotate (hevm_ext.h hevmx_ks_shard_stage): `ranks` target partitions executed stage by stage on ONE
    GPU -- sharing the scratch buffers stands in for the NCCL all-gather / broadcast -- equals the oracle's rotate."""
    from dacapo_b200.sharded import partition_targets
    g, o = pair
    a = o.random_ct(lvl, 900 + lvl)
    o.ct_write(0, a, 2.0 ** 40)
    o.exec(asm.ROTATE, 1, 0, step)
    g.ct_write(0, a, 2.0 ** 40)
    g.ct_write(1, np.zeros_like(a), 2.0 ** 40)
    for stage in (1, 2, 3):
        for tlo, thi in partition_targets(lvl, ranks):
            g.lib.hevmx_ks_shard_stage(g.vm, stage, 1, 0, step, tlo, thi)
    g.lib.hevmx_sync(g.vm)
    assert np.array_equal(g.ct_read(1), o.ct_read(1)), (lvl, step, ranks)


@pytest.mark.parametrize("lvl,ranks,inplace", [(13, 2, False), (6, 3, True), (1, 2, False)])
def test_sharded_mulcc_stages_bit_exact(pair, lvl, ranks, inplace):
    """Limb-sharded multiply + relinearise (hevmx_mulcc_shard_stage) on one GPU, partitions back to back, vs the oracle."""
    from dacapo_b200.sharded import partition_targets
    g, o = pair
    a, b = o.random_ct(lvl, 950 + lvl), o.random_ct(lvl, 960 + lvl)
    for vm in pair:
        vm.ct_write(0, a, 2.0 ** 40)
        vm.ct_write(1, b, 2.0 ** 40)
    dst = 0 if inplace else 2
    o.exec(asm.MULCC, dst, 0, 1)
    if not inplace:
        g.ct_write(2, np.zeros_like(a), 2.0 ** 40)
    for stage in (1, 2, 3):
        for tlo, thi in partition_targets(lvl, ranks):
            g.lib.hevmx_mulcc_shard_stage(g.vm, stage, dst, 0, 1, tlo, thi)
    g.lib.hevmx_sync(g.vm)
    assert np.array_equal(g.ct_read(dst), o.ct_read(dst)), (lvl, ranks, inplace)
    assert g.ct_info(dst) == o.ct_info(dst)


@pytest.mark.parametrize("lvl,n", [(13, 3), (4, 5), (1, 2), (2, 35)])
def test_batched_launch_bit_exact(oracle_lib, b200_lib, tmp_path_factory, lvl, n):
    """hevmx_exec_batch: n independent rotate / mulcc / rescale ops in ONE persistent launch each (ks_fused.cuh; n = 35
    spans two launches of KS_MAX_BATCH = 32) vs the oracle op by op; rotations of one batch use different Galois keys."""
    d = str(tmp_path_factory.mktemp("keysb"))
    g = VM(b200_lib, LOGN, NPR, keydir=d, nct=3 * n, npt=1)
    o = VM(oracle_lib, LOGN, NPR, keydir=d, nct=3 * n, npt=1)
    steps = [1, -1, 64, 2, -8192]
    for k in range(n):
        a = o.random_ct(lvl, 1000 + 7 * k) if k < 6 else None  # six distinct inputs are enough for n = 35
        for vm in (g, o):
            vm.ct_write(k, a if a is not None else vm.ct_read(k % 6), 2.0 ** 40)
    src, dst, dst2 = list(range(n)), list(range(n, 2 * n)), list(range(2 * n, 3 * n))
    rot = [steps[k % len(steps)] for k in range(n)]
    g.exec_batch(asm.ROTATE, dst, src, rot)
    g.exec_batch(asm.MULCC, dst2, dst, src)
    if lvl > 1:
        g.exec_batch(asm.RESCALE, dst2, dst2, [0] * n)  # in place
    for k in range(n):
        o.exec(asm.ROTATE, dst[k], src[k], rot[k])
        o.exec(asm.MULCC, dst2[k], dst[k], src[k])
        if lvl > 1:
            o.exec(asm.RESCALE, dst2[k], dst2[k])
    for k in range(n):
        assert np.array_equal(g.ct_read(dst[k]), o.ct_read(dst[k])), ("rotate", k)
        assert np.array_equal(g.ct_read(dst2[k]), o.ct_read(dst2[k])), ("mulcc+rescale", k)
        assert g.ct_info(dst2[k]) == o.ct_info(dst2[k])


@pytest.mark.parametrize("world", [2, 3])
def test_p2p_sharded_keyswitch_sharded_key_storage(oracle_lib, b200_lib, tmp_path_factory, world):
    """hevmx_ks_shard_p2p with `world` ranks emulated by `world` VMs on ONE GPU in one process: every VM stores only its
    limbs of every key (HEVM_SHARD_*), pushes its digit rows into the other VMs' exchange blocks and waits on their epoch
    flags; three key switches back to back exercise the double-buffered exchange.  Own limbs of every rank, assembled,
    must equal the oracle's rotate / multiply + relinearise."""
    from dacapo_b200.sharded import static_targets
    npr, steps = 7, (1, -2)
    d = str(tmp_path_factory.mktemp(f"keysp2p{world}"))
    o = VM(oracle_lib, LOGN, npr, keydir=d, nct=6, npt=1, galois_steps=steps)
    vms = [VM(b200_lib, LOGN, npr, keydir=d, nct=6, npt=1, galois_steps=steps,
              env={"HEVM_SHARD_RANK": g, "HEVM_SHARD_WORLD": world, "HEVM_STREAMS": 1}) for g in range(world)]
    lib = b200_lib
    whole = o.key(2).size
    assert sum(v.key(2).size for v in vms) == whole and max(v.key(2).size for v in vms) <= whole // world + whole // (npr * 1)
    for g, v in enumerate(vms):
        lib.hevmx_p2p_setup(v.vm, g, world, None)
    for g, v in enumerate(vms):
        for h, w in enumerate(vms):
            if h != g:
                lib.hevmx_p2p_connect(v.vm, h, None, w.vm)
    for lvl in (6, 3, 1):
        a, b = o.random_ct(lvl, 700 + lvl), o.random_ct(lvl, 800 + lvl)
        for vm in [o] + vms:
            vm.ct_write(0, a, 2.0 ** 40)
            vm.ct_write(1, b, 2.0 ** 40)
            vm.ct_write(2, np.zeros_like(a), 2.0 ** 40)
            vm.ct_write(3, np.zeros_like(a), 2.0 ** 40)
        o.exec(asm.ROTATE, 2, 0, 1)
        o.exec(asm.ROTATE, 2, 2, -2)
        o.exec(asm.MULCC, 3, 0, 1)
        # the ranks share one GPU here, so every op is issued phase by phase over all ranks (a rank's wait kernel is then
        # never queued in front of the push it waits for): rotate(1), then rotate(-2) of the result, then mulcc
        def sharded(opcode, dst, lhs, rhs):
            for phase in (1, 2, 3):
                for v in vms:
                    lib.hevmx_ks_shard_p2p_phase(v.vm, phase, opcode, dst, lhs, rhs)

        sharded(1, 2, 0, 1)
        # the second rotation reads register 2, of which every rank only holds ITS limbs: gather them first
        for v in vms:
            lib.hevmx_sync(v.vm)
        parts = [v.ct_read(2) for v in vms]
        full = np.zeros_like(a)
        for g in range(world):
            tlo, thi = static_targets(lvl, npr, g, world)
            full[:, tlo:min(thi, lvl)] = parts[g][:, tlo:min(thi, lvl)]
        for v in vms:
            v.ct_write(2, full, 2.0 ** 40)
        sharded(1, 2, 2, -2 & 0xFFFF)
        sharded(8, 3, 0, 1)
        for v in vms:
            lib.hevmx_sync(v.vm)
        for reg in (2, 3):
            got = np.zeros_like(a)
            for g, v in enumerate(vms):
                tlo, thi = static_targets(lvl, npr, g, world)
                got[:, tlo:min(thi, lvl)] = v.ct_read(reg)[:, tlo:min(thi, lvl)]
            assert np.array_equal(got, o.ct_read(reg)), (lvl, reg)


def test_seal_key_directory_round_trip(b200_lib, tmp_path_factory):
    """dacapo_b200/seal_format.py: VM A's keys are written as a SEAL key directory (parm / pub / sec / relin / gal .seal,
    SEAL_HEVM.cpp:56-88), VM B -- created from ANOTHER seed -- loads them (loadSEAL, 91-129) and from then on computes the
    very same words as A: the path a key set made by a real SEAL 4.0 would take into this backend."""
    from dacapo_b200 import seal_format as sf
    steps = (1, -2, 64)
    a = VM(b200_lib, 14, 5, seed=0xA11CE, keydir=str(tmp_path_factory.mktemp("ka")), nct=4, npt=2, galois_steps=steps)
    b = VM(b200_lib, 14, 5, seed=0xB0B, keydir=str(tmp_path_factory.mktemp("kb")), nct=4, npt=2, galois_steps=(1,))
    assert not np.array_equal(a.key(0), b.key(0))
    d = tmp_path_factory.mktemp("sealdir")
    sf.export_vm_keys(b200_lib, a.vm, d, sf.COMPR_ZSTD)
    kd = sf.load_seal_keys(b200_lib, b.vm, d)
    assert kd["primes"] == a.primes and len(kd["galois"]) == len(steps)
    for which in (0, 1, 2):
        assert np.array_equal(a.key(which), b.key(which))
    x = np.random.default_rng(5).uniform(-1, 1, a.N // 2)
    a.encode(0, x, 4, 50)
    a.encrypt_pt(0, 0, counter=7)
    ct = a.ct_read(0)
    b.ct_write(0, ct, 2.0 ** 50)
    for st in steps:
        e = b200_lib.hevmx_galois_elt(a.vm, st)
        assert e == b200_lib.hevmx_galois_elt(b.vm, st) and np.array_equal(a.key(3, e), b.key(3, e)), ("galois key", st)
    assert np.array_equal(a.ct_read(0), b.ct_read(0))
    for op in ((asm.ROTATE, 1, 0, 64), (asm.ROTATE, 1, 1, -2), (asm.MULCC, 2, 1, 0), (asm.RESCALE, 2, 2, 0)):
        for vm in (a, b):
            vm.exec(*op)
        assert a.ct_info(op[1]) == b.ct_info(op[1]) and np.array_equal(a.ct_read(op[1]), b.ct_read(op[1])), op
    assert np.max(np.abs(b.decrypt_decode(2, 1) - np.roll(x, -62) * x)) < 1e-4   # decrypts under the imported secret key (scale 2^100 / q ~ 2^40)


def test_encode_decode_bit_exact(pair):
    g, o = pair
    rng = np.random.default_rng(7)
    n = o.N // 2
    cases = [(rng.uniform(-1, 1, n), 13, 40), (rng.uniform(-8, 8, 3), 5, 40), (np.array([1.0]), 2, 20),
             (rng.uniform(-1, 1, n), 4, 90), (np.zeros(n), 1, 40), (rng.uniform(-1e3, 1e3, 100), 13, 59)]
    for vals, lvl, sb in cases:
        for vm in pair:
            vm.encode(0, vals, lvl, sb)
        assert np.array_equal(g.pt_read(0), o.pt_read(0)), (lvl, sb)
        assert g.pt_info(0) == o.pt_info(0)
        dg, do = g.decode(0), o.decode(0)
        assert np.array_equal(dg, do), (lvl, sb)
        assert np.max(np.abs(do - np.resize(vals, n))) < 1e-6 * max(1.0, np.max(np.abs(vals)))


def test_encrypt_decrypt_bit_exact(pair):
    g, o = pair
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, o.N // 2)
    for lvl in (13, 3, 1):
        for vm in pair:
            vm.encode(0, x, lvl, 40)
            vm.encrypt_pt(0, 0, counter=1000 + lvl)
        assert np.array_equal(g.ct_read(0), o.ct_read(0)), lvl
        for vm in pair:
            vm.decrypt_to_pt(0, 1)
        assert np.array_equal(g.pt_read(1), o.pt_read(1))
        assert np.max(np.abs(g.decode(1) - x)) < 1e-7


def test_bootstrap_op_bit_exact(pair):
    g, o = pair
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, o.N // 2)
    for vm in pair:
        vm.encode(0, x, 2, 40)
        vm.encrypt_pt(0, 0, counter=77)
        vm.lib.hevmx_set_enc_counter(vm.vm, 78)
        vm.exec(asm.BOOTSTRAP, 1, 0, 13)
    assert np.array_equal(g.ct_read(1), o.ct_read(1))
    assert g.ct_info(1) == o.ct_info(1) == (13, 2.0 ** 40)
    assert np.max(np.abs(g.decrypt_decode(1, 1) - x)) < 1e-6


def _program():
    """A small program touching every opcode the compiler emits, incl. placeholders and aliasing."""
    p = asm.Program(init_level=13)
    x = p.arg(50, 4)
    y = p.arg(50, 4)
    t, u, v = p.new_ct(), p.new_ct(), p.new_ct()
    c = p.const(np.linspace(0.1, 0.9, 37))
    pt, ones, bias = p.new_pt(), p.new_pt(), p.new_pt()
    p.encode(pt, c, 4, 50)
    p.encode(ones, -1, 3, 20)
    p.emit(asm.PLACEHOLDER, 0xBEEF, 0xDEAD, 0xF00D)
    p.emit(asm.MULCP, t, x, pt)          # t = x*w            scale 100, level 4
    p.emit(asm.MULCC, u, x, y)           # u = x*y            scale 100
    p.emit(asm.ADDCC, t, t, u)           # t = x*w + x*y
    p.emit(asm.RESCALE, t, t)            # scale ~40, level 3
    p.encode(bias, p.const([0.25]), 3, 40)
    p.emit(asm.ADDCP, t, t, bias)
    p.rotate(v, t, 5)
    p.rotate(v, v, -12285)
    p.emit(asm.NEGATE, u, v)
    p.emit(asm.ADDCC, v, u, t)           # v = t - rot(t)
    p.emit(asm.MULCP, v, v, ones)        # upscale lowering (UpscaleToMulcp.cpp:52-72)
    p.emit(asm.MODSWITCH, v, v, 1)       # level 2
    p.emit(asm.BOOTSTRAP, u, v, 6)       # decrypt + re-encrypt at level 6
    p.emit(asm.MULCC, u, u, u)
    p.emit(asm.RESCALE, u, u)
    p.result(u, 60, 5)
    p.result(t, 40, 3)
    return p


def test_program_through_c_abi(pair, tmp_path):
    g, o = pair
    p = _program()
    cst, hv = tmp_path / "p.cst", tmp_path / "p.hevm"
    p.save(cst, hv)
    n = o.N // 2
    rng = np.random.default_rng(10)
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    outs = []
    for vm in pair:
        lib = vm.lib
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        lib.hevmx_set_enc_counter(vm.vm, 5)
        assert lib.getArgLen(vm.vm) == 2 and lib.getResLen(vm.vm) == 2
        for i, d in enumerate((x, y)):
            lib.encrypt(vm.vm, i, d.ctypes.data_as(C.POINTER(C.c_double)), n)
        lib.run(vm.vm)
        res = np.zeros((2, n))
        for i in range(2):
            lib.decrypt_result(vm.vm, i, res[i].ctypes.data_as(C.POINTER(C.c_double)))
        outs.append((res, vm.ct_read(lib.getResIdx(vm.vm, 0)), vm.ct_read(lib.getResIdx(vm.vm, 1))))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][0], outs[1][0])
    w = np.resize(np.linspace(0.1, 0.9, 37), n)
    t = x * w + x * y + 0.25
    v = t - np.roll(np.roll(t, -5), 12285)
    assert np.max(np.abs(outs[0][0][1] - t)) < 1e-5
    assert np.max(np.abs(outs[0][0][0] - v * v)) < 1e-4
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


def test_async_encrypt_and_batched_results_bit_exact(pair, tmp_path):
    """encrypt() is asynchronous on the GPU (argument i goes to lane i mod #lanes, vm.cu `encrypt_async`) and
    decrypt_result() decrypts all results at its first call (`decrypt_all_results`): 20 arguments (more than lanes) and
    20 results, two rounds with re-encrypted arguments in between (the cached results must be dropped), an argument
    encrypted twice in a row, and a plain decrypt() in the middle -- ciphertexts and decoded values equal the oracle's."""
    g, o = pair
    nargs = 20
    p = asm.Program(init_level=13)
    xs = [p.arg(40, 2 + (i % 3)) for i in range(nargs)]
    outs = [p.new_ct() for _ in range(nargs)]
    for i in range(nargs):
        p.emit(asm.ADDCC, outs[i], xs[i], xs[i])
        p.result(outs[i], 40, 2 + (i % 3))
    cst, hv = tmp_path / "a.cst", tmp_path / "a.hevm"
    p.save(cst, hv)
    n = o.N // 2
    f64p = C.POINTER(C.c_double)
    got = []
    for vm in pair:
        lib, rng = vm.lib, np.random.default_rng(21)
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        lib.hevmx_set_enc_counter(vm.vm, 1000)
        rounds = []
        for rnd in range(2):
            vals = rng.uniform(-1, 1, (nargs, n))
            buf = np.empty(n)
            for i in range(nargs):
                buf[:] = vals[i]
                lib.encrypt(vm.vm, i, buf.ctypes.data_as(f64p), n)  # the caller's buffer is reused at once
            if rnd == 1:  # argument 3 again, with other values: the second encryption wins
                vals[3] = rng.uniform(-1, 1, n)
                buf[:] = vals[3]
                lib.encrypt(vm.vm, 3, buf.ctypes.data_as(f64p), n)
            buf[:] = 7.0
            lib.run(vm.vm)
            res = np.zeros((nargs, n))
            for i in range(nargs):
                lib.decrypt_result(vm.vm, i, res[i].ctypes.data_as(f64p))
                if i == 5:  # any other entry point in between drops the cached results; the answers do not change
                    one = np.zeros(n)
                    lib.decrypt(vm.vm, lib.getResIdx(vm.vm, 0), one.ctypes.data_as(f64p))
                    assert np.array_equal(one, res[0])
            cts = [vm.ct_read(lib.getResIdx(vm.vm, i)) for i in (0, 3, 17, 19)]
            assert np.max(np.abs(res - 2 * vals)) < 1e-6
            rounds.append((res, cts))
        got.append(rounds)
    for rnd in range(2):
        assert np.array_equal(got[0][rnd][0], got[1][rnd][0])
        for a, b in zip(got[0][rnd][1], got[1][rnd][1]):
            assert np.array_equal(a, b)
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


def test_fused_mulcp_addcc_bit_exact(pair, tmp_path):
    """The scheduler fuses `mulcp t; addcc d <- t + y` (t dead afterwards) into one kernel: every aliasing form, plus the
    forms that must NOT be fused (t read again later), against the oracle's sequential interpreter."""
    g, o = pair
    lv = 3
    p = asm.Program(init_level=13)
    x, y = p.arg(40, lv), p.arg(40, lv)
    regs = [p.new_ct() for _ in range(6)]
    c1, c2 = p.const(np.linspace(-0.9, 0.7, 53)), p.const([0.5, -0.25, 0.125])
    p1, p2 = p.new_pt(), p.new_pt()
    p.encode(p1, c1, lv, 40)
    p.encode(p2, c2, lv, 40)
    t, d, acc, u, w, z = regs
    p.emit(asm.MULCP, t, x, p1); p.emit(asm.ADDCC, d, t, y)        # t as lhs, fresh destination (t dead: overwritten next)
    p.emit(asm.MULCP, t, y, p2); p.emit(asm.ADDCC, t, d, t)        # t as rhs, destination = t
    p.emit(asm.MULCP, acc, x, p2)                                   # plain mulcp (next op is not an addcc of it)
    p.emit(asm.ROTATE, u, acc, 1)
    p.emit(asm.MULCP, w, u, p1); p.emit(asm.ADDCC, acc, acc, w)    # in-place accumulation, w dead (overwritten below)
    p.emit(asm.MULCP, w, w, p2); p.emit(asm.ADDCC, z, w, acc)      # mulcp in place; w is read again below -> not fused
    p.emit(asm.ADDCC, u, w, z)
    p.emit(asm.MULCP, w, x, p1); p.emit(asm.ADDCC, w, w, w)        # both operands are the temporary -> not fused
    for r in (t, d, acc, u, w, z):
        p.result(r, 80, lv)
    cst, hv = tmp_path / "f.cst", tmp_path / "f.hevm"
    p.save(cst, hv)
    n = o.N // 2
    rng = np.random.default_rng(77)
    xs, ys = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    outs = []
    for vm in pair:
        lib = vm.lib
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        lib.hevmx_set_enc_counter(vm.vm, 9)
        for i, dd in enumerate((xs, ys)):
            lib.encrypt(vm.vm, i, dd.ctypes.data_as(C.POINTER(C.c_double)), n)
        lib.run(vm.vm)  # GPU: issued on the lanes
        lib.run(vm.vm)  # GPU: captured into a CUDA graph and launched; inputs are untouched by the program
        lib.run(vm.vm)  # GPU: graph replay
        outs.append([(vm.ct_read(r), vm.ct_info(r)) for r in (t, d, acc, u, w, z)])
    for (a, ia), (b, ib) in zip(*outs):
        assert ia == ib
        assert np.array_equal(a, b)
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


def test_accumulation_chain_bit_exact(pair, tmp_path):
    """acc += rot_k(x) * p_k chains (convolution taps): the scheduler defers the terms and issues one multi-term kernel
    per chain (at most 8 terms); the rotation register is recycled between the taps, a chain is longer than 8, another
    is cut by a read of the accumulator -- all bit-identical to the oracle's sequential interpreter."""
    g, o = pair
    lv = 4
    p = asm.Program(init_level=13)
    x = p.arg(40, lv)
    acc, r, t, z, acc2 = (p.new_ct() for _ in range(5))
    pts = []
    for k in range(11):
        pt = p.new_pt()
        p.encode(pt, p.const(np.linspace(-1 + 0.1 * k, 0.7, 29 + k)), lv, 40)
        pts.append(pt)
    p.emit(asm.MULCP, acc, x, pts[0])                      # acc = x * p0
    for k in range(1, 11):                                  # ten taps: longer than one kernel's 8 terms
        p.rotate(r, x, 1 << (k % 5))                       # the same register r is overwritten for every tap
        p.emit(asm.MULCP, t, r, pts[k])
        p.emit(asm.ADDCC, acc, acc, t)
    p.emit(asm.MULCP, acc2, x, pts[1])
    for k in range(2, 6):
        p.rotate(r, x, -(1 << k))
        p.emit(asm.MULCP, t, r, pts[k])
        p.emit(asm.ADDCC, acc2, t, acc2)                   # accumulator as the right operand
        if k == 3:
            p.emit(asm.ADDCC, z, acc2, acc)                # a read of acc2 in the middle of its chain
    p.emit(asm.MULCP, t, acc2, pts[6])                     # multiplicand is the accumulator itself: no chain
    p.emit(asm.ADDCC, acc2, acc2, t)
    ping, pong = p.new_ct(), p.new_ct()                     # running sum alternates between two registers (SSA-style allocation)
    p.emit(asm.MULCP, ping, x, pts[7])
    cur, nxt = ping, pong
    for k in range(5):
        p.rotate(r, x, 3 + k)                               # composite steps: several key switches each
        p.emit(asm.MULCP, t, r, pts[k])
        p.emit(asm.ADDCC, nxt, cur, t)
        cur, nxt = nxt, cur
    p.emit(asm.ADDCC, z, z, cur)                            # final consumer; `nxt` holds a dead intermediate sum
    p.emit(asm.NEGATE, nxt, x)                              # ... and is overwritten before the program ends
    for reg in (acc, acc2, z, r, cur, nxt):
        p.result(reg, 80, lv)
    cst, hv = tmp_path / "a.cst", tmp_path / "a.hevm"
    p.save(cst, hv)
    n = o.N // 2
    xs = np.random.default_rng(78).uniform(-1, 1, n)
    outs = []
    for vm in pair:
        lib = vm.lib
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        lib.hevmx_set_enc_counter(vm.vm, 11)
        lib.encrypt(vm.vm, 0, xs.ctypes.data_as(C.POINTER(C.c_double)), n)
        for _ in range(3):  # GPU: lanes / graph capture / graph replay
            lib.run(vm.vm)
        outs.append([(vm.ct_read(reg), vm.ct_info(reg)) for reg in (acc, acc2, z, r, cur, nxt)])
    for (a, ia), (b, ib) in zip(*outs):
        assert ia == ib
        assert np.array_equal(a, b)
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


def test_replay_restores_register_metadata(pair, tmp_path):
    """A program that OVERWRITES its argument register and returns it (hecate-opt's ReuseBuffer recycles dead argument
    registers, ReuseBuffer.cpp:27-55): the second run() is a CUDA-graph replay and must leave the result register with
    the level / scale of the first run, not with the argument's (encrypt() in between resets them)."""
    g, o = pair
    p = asm.Program(init_level=13)
    x = p.arg(50, 4)
    t = p.new_ct()
    p.emit(asm.MULCC, t, x, x)
    p.emit(asm.RESCALE, x, t)            # the argument register is recycled: level 3 from here on
    p.rotate(x, x, 3)
    p.emit(asm.MODSWITCH, x, x, 0)       # downFactor 0: a no-op that must not count as a write
    p.result(x, 60, 3)
    cst, hv = tmp_path / "m.cst", tmp_path / "m.hevm"
    p.save(cst, hv)
    n = o.N // 2
    xs = np.random.default_rng(31).uniform(-1, 1, n)
    f64p = C.POINTER(C.c_double)
    outs = {}
    for name, vm in (("g", g), ("o", o)):
        lib = vm.lib
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        for rep in range(3):
            lib.hevmx_set_enc_counter(vm.vm, 40)
            lib.encrypt(vm.vm, 0, xs.ctypes.data_as(f64p), n)
            assert vm.ct_info(0)[0] == 4
            lib.run(vm.vm)
            res = np.zeros(n)
            lib.decrypt_result(vm.vm, 0, res.ctypes.data_as(f64p))
            outs[name, rep] = (vm.ct_info(0), vm.ct_read(0), res)
    for rep in range(3):
        (ig, cg, rg), (io, co, ro) = outs["g", rep], outs["o", rep]
        assert ig == io and ig[0] == 3, rep
        assert np.array_equal(cg, co), rep
        assert np.array_equal(rg, ro), rep
        assert np.max(np.abs(rg - np.roll(xs * xs, -3))) < 1e-5
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


def test_full_size_properties(pair):
    """Size-independent properties on the device alone (no oracle): add/negate cancel exactly,
    rotate(k) o rotate(-k) = id up to noise, rotate and multiply commute with decryption."""
    g, _ = pair
    rng = np.random.default_rng(11)
    n = g.N // 2
    x = rng.uniform(-1, 1, n)
    g.encode(0, x, 13, 40)
    g.encrypt_pt(0, 0)
    g.exec(asm.NEGATE, 1, 0)
    g.exec(asm.ADDCC, 2, 0, 1)
    assert not g.ct_read(2).any()
    g.exec(asm.ROTATE, 3, 0, 4097)
    g.exec(asm.ROTATE, 3, 3, -4097)
    assert np.max(np.abs(g.decrypt_decode(3, 1) - x)) < 1e-4  # slots whose root is near 1 see ~100x the median key-switch noise (non-centred digits)
    y = rng.uniform(-1, 1, n)
    g.encode(1, y, 13, 40)
    g.encrypt_pt(1, 4)
    g.encode(0, x, 13, 50)
    g.encode(1, y, 13, 50)
    g.encrypt_pt(0, 6)
    g.encrypt_pt(1, 7)
    g.exec(asm.ADDCC, 5, 0, 4)
    g.exec(asm.ROTATE, 5, 5, -300)
    assert np.max(np.abs(g.decrypt_decode(5, 1) - np.roll(x + y, 300))) < 1e-4
    g.exec(asm.MULCC, 6, 6, 7)
    g.exec(asm.RESCALE, 6, 6)
    assert np.max(np.abs(g.decrypt_decode(6, 1) - x * y)) < 1e-4  # scale 2^100/q ~ 2^40 after rescale


def _random_program(seed, nops=160):
    """Random dependency-rich program with aggressive register recycling (stress for the multi-lane
    scheduler, register renaming and CUDA-graph replay in run())."""
    rng = np.random.default_rng(seed)
    p = asm.Program(init_level=13)
    args = [p.arg(50, 6), p.arg(50, 6)]
    regs = [p.new_ct() for _ in range(6)]
    lvl = {}
    pts = {}
    for r, a in zip(regs[:2], args):
        p.rotate(r, a, 0)
        lvl[r] = 6
    live = regs[:2]
    for _ in range(nops):
        kind = rng.choice(["rot", "rot", "mulcp", "addcc", "mulcc", "neg", "resc", "addcp", "modsw"])
        src = int(rng.choice(live))
        dst = int(rng.choice(regs))
        l = lvl[src]
        if kind == "rot":
            p.rotate(dst, src, int(rng.choice([1, -2, 64, 96, -4097, 0])))
        elif kind == "neg":
            p.emit(asm.NEGATE, dst, src)
        elif kind in ("mulcp", "addcp"):
            key = (l, kind)
            if key not in pts:
                pts[key] = p.new_pt()
                p.encode(pts[key], p.const(rng.uniform(-1, 1, 5)), l, 30)
            p.emit(asm.MULCP if kind == "mulcp" else asm.ADDCP, dst, src, pts[key])
        elif kind in ("addcc", "mulcc"):
            same = [r for r in live if lvl[r] == l]
            other = int(rng.choice(same))
            p.emit(asm.ADDCC if kind == "addcc" else asm.MULCC, dst, src, other)
        elif kind == "resc":
            if l < 3:
                continue
            p.emit(asm.RESCALE, dst, src)
            l -= 1
        elif kind == "modsw":
            if l < 3:
                continue
            p.emit(asm.MODSWITCH, dst, src, 1)
            l -= 1
        lvl[dst] = l
        if dst not in live:
            live.append(dst)
    for r in live:
        p.result(r, 40, lvl[r])
    return p, live


@pytest.mark.parametrize("seed", [1, 2])
def test_scheduler_random_program_bit_exact(pair, tmp_path, seed):
    g, o = pair
    p, live = _random_program(seed)
    cst, hv = tmp_path / "r.cst", tmp_path / "r.hevm"
    p.save(cst, hv)
    n = o.N // 2
    rng = np.random.default_rng(seed)
    xs = [rng.uniform(-1, 1, n) for _ in range(2)]
    res = {}
    for name, vm in (("g", g), ("o", o)):
        lib = vm.lib
        lib.load(vm.vm, str(cst).encode(), str(hv).encode())
        lib.preprocess(vm.vm)
        lib.hevmx_set_enc_counter(vm.vm, 9)
        for i, d in enumerate(xs):
            lib.encrypt(vm.vm, i, d.ctypes.data_as(C.POINTER(C.c_double)), n)
        lib.run(vm.vm)
        res[name] = [vm.ct_read(lib.getResIdx(vm.vm, i)) for i in range(len(live))]
        if name == "g":  # second run = graph capture + launch, third = replay: from re-encrypted inputs, the same registers
            for rep in ("g2", "g3"):
                lib.hevmx_set_enc_counter(vm.vm, 9)
                for i, d in enumerate(xs):
                    lib.encrypt(vm.vm, i, d.ctypes.data_as(C.POINTER(C.c_double)), n)
                lib.run(vm.vm)
                res[rep] = [vm.ct_read(lib.getResIdx(vm.vm, i)) for i in range(len(live))]
    for a, b, c, d in zip(res["g"], res["o"], res["g2"], res["g3"]):
        assert np.array_equal(a, b)
        assert np.array_equal(c, b)
        assert np.array_equal(d, b)
    for vm in pair:
        vm.lib.hevmx_resize(vm.vm, 8, 4)


@pytest.mark.parametrize("logn,npr", [(14, 5), (16, 6), (17, 4)])
def test_other_ring_sizes_bit_exact(oracle_lib, b200_lib, tmp_path_factory, logn, npr):
    """N is a run-time parameter: same kernels with 64 / 256 rows x 256 columns."""
    d = str(tmp_path_factory.mktemp(f"keys{logn}"))
    g = VM(b200_lib, logn, npr, keydir=d, nct=6, npt=2)
    o = VM(oracle_lib, logn, npr, keydir=d, nct=6, npt=2)
    assert g.primes == o.primes and g.roots == o.roots
    assert np.array_equal(g.key(2), o.key(2))
    lvl = npr - 1
    f = np.random.default_rng(1).integers(0, o.primes[1], size=(2, o.N), dtype=np.uint64)
    assert np.array_equal(g.ntt(f, 1), o.ntt(f, 1))
    a, b = o.random_ct(lvl, 1), o.random_ct(lvl, 2)
    x = np.random.default_rng(3).uniform(-1, 1, o.N // 2)
    for vm in (g, o):
        vm.ct_write(0, a)
        vm.ct_write(1, b)
        vm.exec(asm.MULCC, 2, 0, 1)
        vm.exec(asm.ROTATE, 3, 2, 5)
        vm.exec(asm.ROTATE, 3, 3, -(o.N // 4 + 3))
        vm.exec(asm.RESCALE, 4, 3)
        vm.encode(0, x, lvl, 40)
        vm.encrypt_pt(0, 5, counter=3)
        vm.exec(asm.BOOTSTRAP, 5, 5, lvl)
    for r in (2, 3, 4, 5):
        assert np.array_equal(g.ct_read(r), o.ct_read(r)), r
    assert np.array_equal(g.decrypt_decode(5, 1), o.decrypt_decode(5, 1))
    assert np.max(np.abs(g.decrypt_decode(5, 1) - x)) < 1e-5
