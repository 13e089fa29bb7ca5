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


def bind_near_gpu(local_rank, world=1):
    """Pin this process to the CPUs NVML names as closest to GPU `local_rank`, so that the
    pinned staging buffers it allocates land on that GPU's NUMA node (with 8 ranks feeding
    8 GPUs the host side, not the device, bounds the host-buffer arm).  When NVML names the
    same set for every GPU (one NUMA node), the set is cut into `world` disjoint slices so
    that the ranks' enqueue threads do not share cores.  Returns a short description, or
    None if NVML is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if not cpus:
            return None
        if world > 1:
            try:
                others = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex((local_rank + 1) % world), (ncpu + 63) // 64)
                same = list(others) == list(words)
            except Exception:
                same = False
            if same and len(cpus) >= world:
                per = len(cpus) // world
                cpus = cpus[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cpus near gpu {local_rank} ({cpus[0]}..{cpus[-1]})"
    except Exception:
        return None

