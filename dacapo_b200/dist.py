"""Multi-GPU plumbing for the benchmark: one process per GPU, ranks are independent replicas.

The HEVM path shards only over independent ciphertexts: every rank owns a full VM
(keys replicated) and its own slice of the ciphertext batch; there is NO data-path collective.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for exactly two things: the barrier
around the timed region and the max-over-ranks of the measured time.
"""
import os


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(items, rank, world):
    """Contiguous, balanced slice of `items` for `rank` (first ranks get the remainder)."""
    n = len(items)
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return items[lo:lo + base + (1 if rank < rem else 0)]


def init(backend, device=None):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    kw = {"device_id": device} if device is not None else {}
    dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(x, device="cpu"):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device="cpu"):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput = units processed by all ranks / max over ranks of the elapsed time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
