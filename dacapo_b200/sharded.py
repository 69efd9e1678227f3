"""Limb-sharded key switch (rotation and multiply + relinearise) across the GPUs of one node (BASELINE.json configs[4]).

SEAL's switch_key (reference: the `rotate` opcode, lib/Runtime/SEAL_HEVM.cpp:269-274 -> Evaluator::rotate_vector ->
switch_key_inplace) produces every output limb I from ALL digits J of the operand:

    acc[K][I] = sum_J  NTT_I(INTT_J(c1[J]) mod q_I) * key[J][K][I]

so the natural partition is by OUTPUT limb ("target"): rank g owns a contiguous range of the l+1 targets (data limbs
0..l-1 and the special limb, target l), reads only those limbs of the key and writes only those limbs of the result.
Two exchanges are needed per key switch and they are the only collectives on the path:

    stage 1   own data limbs of perm(c1) -> coefficient digits                 (libB200_HEVM: hevmx_ks_shard_stage 1)
    exchange  all-gather of the digits: l * N * 8 bytes in total                (NCCL over NVLink)
    stage 2   mod-up + key inner product for the own targets; the owner of the
              special limb also runs its inverse NTT and the rounding            (stage 2)
    exchange  broadcast of the two rounded special-limb rows: 2 * N * 8 bytes   (NCCL)
    stage 3   mod-down of the own data limbs                                    (stage 3)

Two implementations of the exchanges:

  * `ShardedRotate` (round 1): NCCL broadcasts issued on the VM's own CUDA stream (torch ExternalStream); every rank
    holds whole keys.  Kept as the comparison arm.
  * `ShardedP2P` (round 2): the library does the exchanges itself over peer memory -- every rank maps the other ranks'
    exchange blocks with CUDA IPC, a push kernel stores its digit rows straight into every peer's HBM over NVLink and
    publishes an epoch flag, a one-warp kernel waits for the peers' flags (`hevmx_ks_shard_p2p`: ONE host call per
    key switch, no NCCL on the data path) -- and the key STORAGE is sharded too: with HEVM_SHARD_RANK / HEVM_SHARD_WORLD
    the VM keeps only its limbs of every key-switch key (28 GB -> 3.5 GB per GPU at N = 2^16, 30 primes, 8 GPUs).

Results are bit-identical to the single-GPU rotate (and to the CPU oracle: tests/test_gpu_parity_l30.py).
"""
from typing import List, Tuple


def partition_targets(level: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ranges [tlo, thi) of the level+1 key-switch targets; the last rank owns the special limb."""
    n = level + 1
    return [(n * g // world, n * (g + 1) // world) for g in range(world)]


def digit_exchange_plan(level: int, world: int) -> List[Tuple[int, int, int]]:
    """(owner rank, first row, end row) of the coefficient-digit rows each rank contributes to the all-gather."""
    plan = []
    for g, (tlo, thi) in enumerate(partition_targets(level, world)):
        hi = min(thi, level)
        if hi > tlo:
            plan.append((g, tlo, hi))
    return plan


def special_owner(level: int, world: int) -> int:
    for g, (_, thi) in enumerate(partition_targets(level, world)):
        if thi == level + 1:
            return g
    raise AssertionError


class _DevMem:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface v2, int64 words)."""

    def __init__(self, ptr: int, nwords: int):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class ShardedRotate:
    """One instance per rank (one process per GPU).  `lib`/`vm`: a bound libB200_HEVM.so and its VM on this rank's GPU."""

    def __init__(self, lib, vm, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.lib, self.vm, self.rank, self.world, self.group = lib, vm, rank, world, group
        self.N = 1 << lib.hevmx_param(vm, 0)
        self.L = lib.hevmx_param(vm, 1)
        self.stream = torch.cuda.ExternalStream(lib.hevmx_stream(vm))
        self.digits = torch.as_tensor(_DevMem(lib.hevmx_dev_ptr(vm, 0), self.L * self.N), device="cuda").view(self.L, self.N)
        self.rnd = torch.as_tensor(_DevMem(lib.hevmx_dev_ptr(vm, 1), 2 * self.N), device="cuda").view(2, self.N)

    def ct_tensor(self, reg: int):
        """[2][L-1][N] int64 view of a ciphertext register's limb storage."""
        t = self.torch.as_tensor(_DevMem(self.lib.hevmx_dev_ptr(self.vm, 16 + reg), 2 * (self.L - 1) * self.N), device="cuda")
        return t.view(2, self.L - 1, self.N)

    def rotate(self, dst: int, src: int, step: int, level: int):
        """dst <- rotate(src, step) with the key switch sharded over the ranks; every rank ends up holding the limbs
        it owns in register `dst` (use `gather` to replicate them)."""
        lib, vm, dist = self.lib, self.vm, self.dist
        tlo, thi = partition_targets(level, self.world)[self.rank]
        with self.torch.cuda.stream(self.stream):
            lib.hevmx_ks_shard_stage(vm, 1, dst, src, step, tlo, thi)
            for owner, a, b in digit_exchange_plan(level, self.world):
                dist.broadcast(self.digits[a:b], src=owner, group=self.group)
            lib.hevmx_ks_shard_stage(vm, 2, dst, src, step, tlo, thi)
            dist.broadcast(self.rnd, src=special_owner(level, self.world), group=self.group)
            lib.hevmx_ks_shard_stage(vm, 3, dst, src, step, tlo, thi)

    def mulcc(self, dst: int, lhs: int, rhs: int, level: int):
        """dst <- lhs * rhs with relinearisation, sharded the same way (hevmx_mulcc_shard_stage); dst may alias an operand."""
        lib, vm, dist = self.lib, self.vm, self.dist
        tlo, thi = partition_targets(level, self.world)[self.rank]
        with self.torch.cuda.stream(self.stream):
            lib.hevmx_mulcc_shard_stage(vm, 1, dst, lhs, rhs, tlo, thi)
            for owner, a, b in digit_exchange_plan(level, self.world):
                dist.broadcast(self.digits[a:b], src=owner, group=self.group)
            lib.hevmx_mulcc_shard_stage(vm, 2, dst, lhs, rhs, tlo, thi)
            dist.broadcast(self.rnd, src=special_owner(level, self.world), group=self.group)
            lib.hevmx_mulcc_shard_stage(vm, 3, dst, lhs, rhs, tlo, thi)

    def gather(self, reg: int, level: int):
        """Replicate a limb-sharded register: every owner broadcasts its data limbs of both polynomials."""
        ct = self.ct_tensor(reg)
        with self.torch.cuda.stream(self.stream):
            for owner, a, b in digit_exchange_plan(level, self.world):
                for k in range(2):
                    self.dist.broadcast(ct[k, a:b], src=owner, group=self.group)


def static_targets(level: int, nprimes: int, rank: int, world: int) -> Tuple[int, int]:
    """Round-2 ownership: rank g owns a contiguous, balanced range of the L = nprimes limbs of the key level (the special
    limb L-1 belongs to the last rank) at EVERY level, so that key storage can be sharded once; at `level` the owned key-switch
    targets are that range cut to [0, level) plus, for the last rank, the special target `level`."""
    base, rem = divmod(nprimes, world)  # the remainder goes to the FIRST ranks (VM::own_range): the special limb's owner is never the busiest
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return min(lo, level), (level + 1 if hi == nprimes else min(hi, level))


class ShardedP2P:
    """One instance per rank.  `vm` must have been created with HEVM_SHARD_RANK / HEVM_SHARD_WORLD in the environment for
    sharded key storage (or without, for whole keys and sharded work only)."""

    def __init__(self, lib, vm, rank: int, world: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.lib, self.vm, self.rank, self.world = lib, vm, rank, world
        handle = C.create_string_buffer(64)
        lib.hevmx_p2p_setup(vm, rank, world, handle)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle.raw, group=group)
            for g in range(world):
                if g != rank:
                    lib.hevmx_p2p_connect(vm, g, handles[g], None)
            dist.barrier(group=group)

    def rotate(self, dst: int, src: int, step: int):
        self.lib.hevmx_ks_shard_p2p(self.vm, 1, dst, src, step & 0xFFFF)

    def mulcc(self, dst: int, lhs: int, rhs: int):
        self.lib.hevmx_ks_shard_p2p(self.vm, 8, dst, lhs, rhs)

    def timing(self, fn):
        """Per-stage CUDA-event breakdown (microseconds) of one sharded op issued by `fn`."""
        import ctypes as C
        out = (C.c_double * 5)()
        self.lib.hevmx_p2p_timing(self.vm, 1, out)
        fn()
        self.lib.hevmx_sync(self.vm)
        self.lib.hevmx_p2p_timing(self.vm, 0, out)
        return {k: round(v * 1e3, 1) for k, v in zip(("stage1", "digit_push", "stage2_with_digit_waits", "row_exchange", "stage3"), out)}


def measure(lib, rank: int, world: int, logn: int = 16, nprimes: int = 30, levels=None, reps: int = 20, seed: int = 0xDACA90):
    """Bit-exactness and device-time of the sharded rotate against the single-GPU rotate on this node's GPUs.
    Every rank derives the same keys from `seed`.  Returns a dict (identical on all ranks).  Needs torch.distributed
    (NCCL) initialised and the current CUDA device set."""
    import ctypes as C
    import os
    import tempfile

    import numpy as np
    import torch
    import torch.distributed as dist

    from . import hevm_asm as asm
    keydir = tempfile.mkdtemp(prefix=f"hevm_shard_keys_r{rank}_")
    old = {k: os.environ.get(k) for k in ("HEVM_LOGN", "HEVM_NUM_PRIMES", "HEVM_SEED", "HEVM_PRIME_BITS")}
    os.environ.update(HEVM_LOGN=str(logn), HEVM_NUM_PRIMES=str(nprimes), HEVM_SEED=str(seed), HEVM_PRIME_BITS="60")
    try:
        lib.create_context(keydir.encode())
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    vm = lib.initFullVM(keydir.encode(), True)
    lib.hevmx_resize(vm, 4, 1)
    N, u64p = 1 << logn, C.POINTER(C.c_uint64)
    primes = np.zeros(nprimes, dtype=np.uint64)
    lib.hevmx_primes(vm, primes.ctypes.data_as(u64p))
    sr = ShardedRotate(lib, vm, rank, world)
    step = 4
    out = {"ring": f"N=2^{logn}, {nprimes} x 60-bit primes", "world": world, "rotate_step": step, "levels": {}}
    # round-2 arm: a second VM whose key STORAGE is sharded (HEVM_SHARD_*), exchanges over peer memory
    os.environ.update(HEVM_SHARD_RANK=str(rank), HEVM_SHARD_WORLD=str(world))
    try:
        vm2 = lib.initFullVM(keydir.encode(), True)
    finally:
        os.environ.pop("HEVM_SHARD_RANK", None), os.environ.pop("HEVM_SHARD_WORLD", None)
    lib.hevmx_resize(vm2, 4, 1)
    p2p = ShardedP2P(lib, vm2, rank, world)
    kb = lambda v, which: lib.hevmx_key_read(v, which, 0, None) * 8
    out["key_bytes_per_rank"] = {"whole_relin_key": kb(vm, 2), "sharded_relin_key": kb(vm2, 2), "keys_per_vm": 1 + lib.hevmx_param(vm, 5)}

    def timed(fn):
        for _ in range(3):
            fn()
        lib.hevmx_sync(vm)
        dist.barrier()
        torch.cuda.synchronize()
        lib.hevmx_timer(vm, 0)
        for _ in range(reps):
            fn()
        t = torch.tensor([lib.hevmx_timer(vm, 1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3

    # bandwidth of the exchange path itself: every rank pushes `words` u64 into all peers at once (all-to-all over NVSwitch)
    bw = {}
    for words in (4 * N, nprimes * N):
        dist.barrier()
        torch.cuda.synchronize()
        ms = lib.hevmx_p2p_bench(vm2, words, 10)
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bw[f"{words * 8 / 1e6:.1f}MB_per_peer"] = {"us": round(float(t.item()) * 1e3, 1),
                                                  "egress_GBps_per_gpu": round(words * 8 * (world - 1) / (float(t.item()) * 1e-3) / 1e9, 1)}
    out["p2p_push_bandwidth"] = bw
    for lvl in levels or (nprimes - 1, (nprimes - 1) // 2):
        rng = np.random.default_rng(1000 + lvl)  # same ciphertext on every rank
        a = np.zeros((2, lvl, N), dtype=np.uint64)
        for i in range(lvl):
            a[:, i, :] = rng.integers(0, int(primes[i]), size=(2, N), dtype=np.uint64)
        lib.hevmx_ct_write(vm, 0, a.ctypes.data_as(u64p), lvl, 2.0 ** 40)
        lib.hevmx_ct_write(vm, 2, np.zeros_like(a).ctypes.data_as(u64p), lvl, 2.0 ** 40)
        lib.hevmx_exec(vm, asm.ROTATE, 1, 0, step)
        exp = np.zeros_like(a)
        lib.hevmx_ct_read(vm, 1, exp.ctypes.data_as(u64p))
        sr.rotate(2, 0, step, lvl)
        sr.gather(2, lvl)
        lib.hevmx_sync(vm)
        got = np.zeros_like(a)
        lib.hevmx_ct_read(vm, 2, got.ctypes.data_as(u64p))
        ok = torch.tensor([1 if np.array_equal(got, exp) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        t_sh = timed(lambda: sr.rotate(2, 0, step, lvl))
        t_1 = timed(lambda: lib.hevmx_exec(vm, asm.ROTATE, 1, 0, step))
        # multiply + relinearise, same protocol
        lib.hevmx_exec(vm, asm.MULCC, 1, 0, 0)
        lib.hevmx_ct_read(vm, 1, exp.ctypes.data_as(u64p))
        lib.hevmx_ct_write(vm, 3, np.zeros_like(a).ctypes.data_as(u64p), lvl, 2.0 ** 40)
        sr.mulcc(3, 0, 0, lvl)
        sr.gather(3, lvl)
        lib.hevmx_sync(vm)
        lib.hevmx_ct_read(vm, 3, got.ctypes.data_as(u64p))
        ok2 = torch.tensor([1 if np.array_equal(got, exp) else 0], device="cuda")
        dist.all_reduce(ok2, op=dist.ReduceOp.MIN)
        m_sh = timed(lambda: sr.mulcc(3, 0, 0, lvl))
        m_1 = timed(lambda: lib.hevmx_exec(vm, asm.MULCC, 1, 0, 0))
        # ---- peer-memory arm (sharded key storage): own limbs of the result vs the single-GPU result
        tlo, thi = static_targets(lvl, nprimes, rank, world)
        dhi = min(thi, lvl)

        def timed2(fn):
            for _ in range(3):
                fn()
            lib.hevmx_sync(vm2)
            dist.barrier()
            torch.cuda.synchronize()
            lib.hevmx_timer(vm2, 0)
            for _ in range(reps):
                fn()
            t = torch.tensor([lib.hevmx_timer(vm2, 1) / reps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) * 1e3

        def own_limbs_equal(reg, ref):
            g2 = np.zeros_like(a)
            lib.hevmx_sync(vm2)
            lib.hevmx_ct_read(vm2, reg, g2.ctypes.data_as(u64p))
            okk = torch.tensor([1 if np.array_equal(g2[:, tlo:dhi], ref[:, tlo:dhi]) else 0], device="cuda")
            dist.all_reduce(okk, op=dist.ReduceOp.MIN)
            return bool(okk.item())

        lib.hevmx_ct_write(vm2, 0, a.ctypes.data_as(u64p), lvl, 2.0 ** 40)
        lib.hevmx_ct_write(vm2, 2, np.zeros_like(a).ctypes.data_as(u64p), lvl, 2.0 ** 40)
        lib.hevmx_ct_write(vm2, 3, np.zeros_like(a).ctypes.data_as(u64p), lvl, 2.0 ** 40)
        lib.hevmx_exec(vm, asm.ROTATE, 1, 0, step)
        lib.hevmx_ct_read(vm, 1, exp.ctypes.data_as(u64p))
        p2p.rotate(2, 0, step)
        ok3 = own_limbs_equal(2, exp)
        lib.hevmx_exec(vm, asm.MULCC, 1, 0, 0)
        lib.hevmx_ct_read(vm, 1, exp.ctypes.data_as(u64p))
        p2p.mulcc(3, 0, 0)
        ok4 = own_limbs_equal(3, exp)
        t_p2p = timed2(lambda: p2p.rotate(2, 0, step))
        m_p2p = timed2(lambda: p2p.mulcc(3, 0, 0))
        dist.barrier()
        stages = p2p.timing(lambda: p2p.rotate(2, 0, step))
        out["levels"][str(lvl)] = {"bit_exact_vs_single_gpu": bool(ok.item()) and bool(ok2.item()) and ok3 and ok4,
                                   "rotate": {"sharded_us": round(t_p2p, 1), "single_gpu_us": round(t_1, 1), "speedup": round(t_1 / t_p2p, 3),
                                              "nccl_broadcast_arm_us": round(t_sh, 1), "nccl_broadcast_arm_speedup": round(t_1 / t_sh, 3),
                                              "stages_us_rank0": stages},
                                   "mulcc": {"sharded_us": round(m_p2p, 1), "single_gpu_us": round(m_1, 1), "speedup": round(m_1 / m_p2p, 3),
                                             "nccl_broadcast_arm_us": round(m_sh, 1), "nccl_broadcast_arm_speedup": round(m_1 / m_sh, 3)},
                                   "targets_per_rank_p2p": [static_targets(lvl, nprimes, g, world)[1] - static_targets(lvl, nprimes, g, world)[0] for g in range(world)],
                                   "exchange": "peer-memory push over NVLink + epoch flags (CUDA IPC), sharded key storage",
                                   "targets_per_rank": [b - a_ for a_, b in partition_targets(lvl, world)],
                                   "allgather_bytes": lvl * N * 8, "broadcast_bytes": 2 * N * 8}
    return out
