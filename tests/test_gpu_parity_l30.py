"""GPU tier (-m gpu): the long modulus chain of BASELINE.json configs[4] -- N = 2^16, 30 x 60-bit primes -- against the
CPU oracle, bit-exact.  Levels above MAC_FLUSH_DIGITS = 15 take the second accumulation chunk of the key inner product
(ntt_bodies.cuh body_mac_dot), which no N = 2^15 / 14-prime test can reach; the limb-sharded stages
(hevmx_ks_shard_stage / hevmx_mulcc_shard_stage, the 8-GPU path of dacapo_b200/sharded.py) and the batched single-launch
form (hevmx_exec_batch) are checked at the same levels.  Only three key-switch keys are generated
(HEVM_GALOIS_STEPS = SEAL's create_galois_keys(steps)): one key is 0.9 GB at this geometry."""
import numpy as np
import pytest

from dacapo_b200 import hevm_asm as asm
from dacapo_b200.sharded import partition_targets
from util import VM

pytestmark = pytest.mark.gpu
LOGN, NPR, STEPS = 16, 30, (1, -2)


@pytest.fixture(scope="module")
def pair(oracle_lib, b200_lib, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("keys30"))
    g = VM(b200_lib, LOGN, NPR, keydir=d, nct=8, npt=1, galois_steps=STEPS)
    o = VM(oracle_lib, LOGN, NPR, keydir=d, nct=8, npt=1, galois_steps=STEPS)
    assert g.primes == o.primes and len(g.primes) == 30
    return g, o


def test_keys_bit_exact_l30(pair):
    g, o = pair
    assert np.array_equal(g.key(2), o.key(2))
    for st in STEPS:
        elt = o.lib.hevmx_galois_elt(o.vm, st)
        assert np.array_equal(g.key(3, elt), o.key(3, elt)), st
    assert g.lib.hevmx_param(g.vm, 5) == o.lib.hevmx_param(o.vm, 5) == len(STEPS)


@pytest.mark.parametrize("lvl", [16, 29])
def test_rotate_mulcc_rescale_above_15_limbs(pair, lvl):
    g, o = pair
    a, b = o.random_ct(lvl, 3000 + lvl), o.random_ct(lvl, 3100 + lvl)
    for vm in pair:
        vm.ct_write(0, a, 2.0 ** 40)
        vm.ct_write(1, b, 2.0 ** 40)
        vm.exec(asm.ROTATE, 2, 0, 1)
        vm.exec(asm.ROTATE, 3, 1, -2)
        vm.exec(asm.MULCC, 4, 2, 3)
        vm.exec(asm.MULCC, 5, 0, 0)
        vm.exec(asm.RESCALE, 6, 4)
        vm.exec(asm.RESCALE, 5, 5)
    for r in (2, 3, 4, 5, 6):
        assert np.array_equal(g.ct_read(r), o.ct_read(r)), (lvl, r)
        assert g.ct_info(r) == o.ct_info(r)


@pytest.mark.parametrize("lvl,ranks", [(29, 8), (16, 3)])
def test_sharded_stages_above_15_limbs(pair, lvl, ranks):
    """The limb-sharded rotate and multiply+relinearise of the 8-GPU path, partitions run back to back on one GPU."""
    g, o = pair
    a, b = o.random_ct(lvl, 3200 + lvl), o.random_ct(lvl, 3300 + lvl)
    for vm in pair:
        vm.ct_write(0, a, 2.0 ** 40)
        vm.ct_write(1, b, 2.0 ** 40)
    o.exec(asm.ROTATE, 2, 0, 1)
    o.exec(asm.MULCC, 3, 0, 1)
    g.ct_write(2, np.zeros_like(a), 2.0 ** 40)
    g.ct_write(3, np.zeros_like(a), 2.0 ** 40)
    parts = partition_targets(lvl, ranks)
    for stage in (1, 2, 3):
        for tlo, thi in parts:
            g.lib.hevmx_ks_shard_stage(g.vm, stage, 2, 0, 1, tlo, thi)
    for stage in (1, 2, 3):
        for tlo, thi in parts:
            g.lib.hevmx_mulcc_shard_stage(g.vm, stage, 3, 0, 1, tlo, thi)
    g.lib.hevmx_sync(g.vm)
    assert np.array_equal(g.ct_read(2), o.ct_read(2)), "sharded rotate"
    assert np.array_equal(g.ct_read(3), o.ct_read(3)), "sharded mulcc"
    assert g.ct_info(3) == o.ct_info(3)


def test_batched_launch_above_15_limbs(pair):
    g, o = pair
    lvl, n = 17, 2
    for k in range(n):
        a = o.random_ct(lvl, 3400 + k)
        for vm in pair:
            vm.ct_write(k, a, 2.0 ** 40)
    g.exec_batch(asm.ROTATE, [2, 3], [0, 1], [1, -2])
    g.exec_batch(asm.MULCC, [4, 5], [2, 3], [0, 1])
    g.exec_batch(asm.RESCALE, [4, 5], [4, 5], [0, 0])
    for k in range(n):
        o.exec(asm.ROTATE, 2 + k, k, (1, -2)[k])
        o.exec(asm.MULCC, 4 + k, 2 + k, k)
        o.exec(asm.RESCALE, 4 + k, 4 + k)
    for r in (2, 3, 4, 5):
        assert np.array_equal(g.ct_read(r), o.ct_read(r)), r
