"""CPU tier: the N>1 plumbing of bench.py (replica sharding, barrier, max-over-ranks) under gloo, world_size 2."""
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys, time
sys.path.insert(0, os.environ["REPO"])
from dacapo_b200 import dist as D
rank, world, _ = D.rank_info()
dist = D.init("gloo")
assert dist.get_world_size() == 2 == world
cts = list(range(7))
mine = D.shard(cts, rank, world)
assert mine == ([0, 1, 2, 3] if rank == 0 else [4, 5, 6])
dist.barrier()
secs = 1.0 + rank            # rank 1 is slower
agg = D.aggregate_throughput(len(mine) * 38, secs)
assert abs(agg - 7 * 38 / 2.0) < 1e-9, agg
assert D.max_over_ranks(secs) == 2.0
dist.barrier()
dist.destroy_process_group()
os.write(1, f"rank{rank}ok\n".encode())
"""


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_static_target_ownership():
    """Round-2 static ownership of key-switch targets (sharded.static_targets == VM::own_range cut to a level): a partition
    of [0, level], balanced to within one, the special target with the last rank, and the last rank never the busiest."""
    sys.path.insert(0, str(REPO))
    from dacapo_b200.sharded import static_targets
    for L in (14, 30, 7):
        for w in (1, 2, 3, 8):
            if w > L // 2:
                continue  # every rank owns at least two limbs (the VM refuses anything else)
            full = [static_targets(L - 1, L, g, w) for g in range(w)]
            sizes = [hi - lo for lo, hi in full]
            assert full[0][0] == 0 and full[-1][1] == L and all(full[g][1] == full[g + 1][0] for g in range(w - 1))
            assert max(sizes) - min(sizes) <= 1 and sizes[-1] == min(sizes)
            for lvl in range(1, L):
                parts = [static_targets(lvl, L, g, w) for g in range(w)]
                cover = sorted(t for lo, hi in parts for t in range(lo, hi))
                assert cover == list(range(lvl + 1)), (L, w, lvl, parts)
                assert parts[-1][1] == lvl + 1


def test_shard_balanced():
    sys.path.insert(0, str(REPO))
    from dacapo_b200 import dist as D
    for n in (0, 1, 5, 28):
        for w in (1, 2, 4, 8):
            parts = [D.shard(list(range(n)), r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


SHARD_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["REPO"])
import torch, torch.distributed as dist
from dacapo_b200.sharded import partition_targets, digit_exchange_plan, special_owner
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
level, N = 13, 64
tlo, thi = partition_targets(level, world)[rank]
# the exchange wiring of ShardedRotate.rotate on CPU tensors: every rank fills the digit rows it owns ...
digits = torch.full((level + 1, N), -1, dtype=torch.int64)
digits[tlo:min(thi, level)] = rank
for owner, a, b in digit_exchange_plan(level, world):
    dist.broadcast(digits[a:b], src=owner)
# ... and afterwards every rank sees every digit row tagged by its owner
for g, (a, b) in enumerate(partition_targets(level, world)):
    assert (digits[a:min(b, level)] == g).all()
rnd = torch.full((2, N), rank, dtype=torch.int64)
dist.broadcast(rnd, src=special_owner(level, world))
assert (rnd == world - 1).all()
dist.barrier()
dist.destroy_process_group()
os.write(1, f"shard{rank}ok\n".encode())
"""


def test_sharded_keyswitch_exchange_plan_gloo(tmp_path):
    """world_size-2 gloo run of the two exchanges of the limb-sharded key switch (all-gather of digits, broadcast of the rounding rows)."""
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER)
    env = dict(os.environ, REPO=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shard0ok" in r.stdout and "shard1ok" in r.stdout


def test_target_partition_properties():
    sys.path.insert(0, str(REPO))
    from dacapo_b200.sharded import partition_targets, digit_exchange_plan, special_owner
    for level in (1, 2, 7, 13, 29):
        for w in (1, 2, 3, 4, 8):
            parts = partition_targets(level, w)
            assert parts[0][0] == 0 and parts[-1][1] == level + 1
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
            assert special_owner(level, w) == max(g for g, (a, b) in enumerate(parts) if b > a)
            rows = sorted(r for _, a, b in digit_exchange_plan(level, w) for r in range(a, b))
            assert rows == list(range(level))
