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


def test_shard_balanced():
    sys.path.insert(0, str(REPO))
    from dacapo_b200 import dist as D
    for n in (0, 1, 5, 28):
        for w in (1, 2, 4, 8):
            parts = [D.shard(list(range(n)), r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
