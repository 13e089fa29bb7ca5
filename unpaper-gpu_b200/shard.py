"""Page sharding across GPUs (one process per GPU; no collective on the data
path — sheets are independent, reference sheet_process.h:20-21).

torch.distributed is used only for the barrier around a timed region and for
the MAX over ranks of the measured time."""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_total, rank, world):
    """Contiguous, balanced [lo, hi) slice of job indices for `rank`
    (what `dev = job % G` does in spirit, kept contiguous for locality)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]
