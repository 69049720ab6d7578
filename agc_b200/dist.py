"""torch.distributed plumbing for one-process-per-GPU runs.

Two ways to use N GPUs:
  * replicas / sample shards without a data-path collective (`shard_samples`; what bench.py --gpus N measures: weak scaling);
  * ONE archive from N GPUs (`install_exchange`): every rank makes the same CAGCCompressor calls on the same inputs; the host
    bookkeeping and the cheap device passes are replicated, LZ-diff encoding and the residual coder -- >90 % of the device time
    -- are split across the ranks and their results all-gathered (NCCL over NVLink on a GPU box, gloo in the CPU tests).  Rank 0
    writes the archive, byte-identical to the single-GPU one (include/agcgpu.h, agcgpu_set_exchange).
The other collectives here are the barrier / max-over-ranks used for timing."""
import ctypes as C
import os

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
_keep = []      # the ctypes callback must outlive every compressor that uses it


def env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def install_exchange(L, device=None):
    """Register torch.distributed's all-gather as the exchange step of library `L` (agcgpu_set_exchange).  The process group
    must be initialised.  device=None: CPU tensors (gloo); device="cuda:i": blocks are staged through HBM and gathered with
    NCCL (all_gather_into_tensor).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    L.agcgpu_set_exchange.restype = C.c_int
    L.agcgpu_set_exchange.argtypes = [C.c_uint32, C.c_uint32, ALLGATHER_FN, C.c_void_p]

    def allgather(_user, send, recv, nbytes):
        try:
            src = torch.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=torch.uint8)
            dst = torch.frombuffer((C.c_uint8 * (nbytes * world)).from_address(recv), dtype=torch.uint8)
            if device is None:
                dist.all_gather_into_tensor(dst, src)
            else:
                d_dst = torch.empty(nbytes * world, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(d_dst, src.to(device, non_blocking=True))
                dst.copy_(d_dst)
            return 0
        except Exception as e:          # no exception may cross the C boundary
            print(f"agc_b200.dist: all-gather failed: {e!r}", flush=True)
            return -1

    cb = ALLGATHER_FN(allgather)
    _keep.append(cb)
    rc = L.agcgpu_set_exchange(rank, world, cb, None)
    if rc != 0:
        raise RuntimeError(f"agcgpu_set_exchange failed ({rc})")
    return rank, world


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
