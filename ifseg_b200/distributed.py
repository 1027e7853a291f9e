"""Data-parallel plumbing of the benchmark/serving path: one process per GPU, image-sharded,
no data-path collective (inference = independent replicas; training = fairseq DDP
gradient all-reduce, which stays torch.nn.parallel.DistributedDataParallel over NCCL).
Only the timing reduction (max over ranks) and the shard bookkeeping live here."""
import os

import torch
import torch.distributed as dist


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_seed(base_seed: int, rank: int) -> int:
    """Every rank draws its own synthetic batch (FileDataset shards rows by rank the same way,
    data/file_dataset.py:97-103)."""
    return base_seed + rank


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n_items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (the slowest rank defines the step)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_throughput(units_per_rank: int, world: int, seconds_max_over_ranks: float) -> float:
    """Whole-job units/s = units all ranks processed / slowest rank's time."""
    return units_per_rank * world / seconds_max_over_ranks
