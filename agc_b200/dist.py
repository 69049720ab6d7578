"""torch.distributed plumbing for one-process-per-GPU runs: samples shard across ranks without a data-path collective;
the only collectives are the barrier / max-over-ranks used for timing and the gather of per-rank totals."""
import os


def env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_samples(n_samples, rank, world):
    """contiguous, balanced shard of sample indices 1..n_samples-1 (sample 0 = the reference, which every rank ingests to
    derive the identical splitter set and reference segments)."""
    rest = list(range(1, n_samples))
    per, extra = divmod(len(rest), world)
    lo = rank * per + min(rank, extra)
    hi = lo + per + (1 if rank < extra else 0)
    return [0] + rest[lo:hi]


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
