"""Limb-sharded rotation across GPUs: bit-exactness against the single-GPU rotate, and timing.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_probe.py [logN] [L]"""
import ctypes as C, json, os, sys, tempfile
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch, torch.distributed as dist
from dacapo_b200 import _binding, hevm_asm as asm
from dacapo_b200.sharded import ShardedRotate, partition_targets
from util import VM

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 15
L = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["HEVM_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl")
lib = _binding.bind(_binding.B200_LIB)
g = VM(lib, logn, L, keydir=tempfile.mkdtemp(), nct=6, npt=1)  # same seed on every rank -> same keys
sr = ShardedRotate(lib, g.vm, rank, world)
out = {"logN": logn, "L": L, "world": world}
for lvl in sorted({L - 1, max(2, (L - 1) // 2)}, reverse=True):
    a = g.random_ct(lvl, 5)
    g.ct_write(0, a)
    g.exec(asm.ROTATE, 1, 0, 4)          # single-GPU reference on this rank
    exp = g.ct_read(1)
    g.ct_write(2, np.zeros_like(a))
    sr.rotate(2, 0, 4, lvl); sr.gather(2, lvl); lib.hevmx_sync(g.vm)
    got = g.ct_read(2)
    ok = bool(np.array_equal(got, exp))
    flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # timing: device time on the VM stream, max over ranks
    def timed(fn, reps=20):
        for _ in range(3): fn()
        lib.hevmx_sync(g.vm); dist.barrier(); torch.cuda.synchronize()
        lib.hevmx_timer(g.vm, 0)
        for _ in range(reps): fn()
        ms = lib.hevmx_timer(g.vm, 1) / reps
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3
    t_shard = timed(lambda: sr.rotate(2, 0, 4, lvl))
    t_single = timed(lambda: lib.hevmx_exec(g.vm, asm.ROTATE, 1, 0, 4))
    out["level_%d" % lvl] = {"bit_exact": bool(flag.item()), "sharded_us": round(t_shard, 1), "single_gpu_us": round(t_single, 1),
                             "targets": partition_targets(lvl, world)}
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
