"""Batch-shard data parallelism of the hot path.

Every kernel of the path is per-utterance, so N GPUs run N independent shards of the batch with
replicated weights and NO data-path collective.  What does cross ranks is the same thing the reference
exchanges (`aps/distributed/backend.py:118-129`: a scalar all-reduce): a tiny vector of metrics.
One process per GPU (`torchrun`), backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
import os
from typing import Dict, Tuple

import torch as th
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults otherwise)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Utterances [lo, hi) owned by `rank`: contiguous, sizes differ by at most one, every item owned once
    (the reference gives each rank `batch_size // num_process` utterances, aps/libs.py:265)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(num_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_metrics(local: Dict[str, float], maxima=("elapsed_ms",), device="cpu") -> Dict[str, float]:
    """All-reduce a small dict of floats: keys in `maxima` by MAX (timings: the slowest rank defines the
    step), everything else by SUM (frames, utterances, loss*n).  A no-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    keys = sorted(local)
    mx = th.tensor([local[k] if k in maxima else float("-inf") for k in keys], dtype=th.float64, device=device)
    sm = th.tensor([0.0 if k in maxima else local[k] for k in keys], dtype=th.float64, device=device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return {k: float(mx[i]) if k in maxima else float(sm[i]) for i, k in enumerate(keys)}
