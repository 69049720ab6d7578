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


def install_exchange(device_index):
    """GPU runs: create the NCCL communicator INSIDE libagcgpu (agc_b200/csrc/comm.cu): rank 0 asks the library for an ncclUniqueId,
    torch.distributed (already initialised, any backend) carries the 128 bytes to the other ranks, every rank calls
    agcgpu_comm_init.  From then on compressors of this process split LZ-diff encoding and residual coding over the ranks and
    all-gather the results between device buffers from C++ -- no Python on the data path.  Returns a dict for bench.py."""
    import torch
    import torch.distributed as dist
    from . import lib
    L = lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    L.agcgpu_comm_unique_id.restype = C.c_int; L.agcgpu_comm_unique_id.argtypes = [C.c_void_p]
    L.agcgpu_comm_init.restype = C.c_int; L.agcgpu_comm_init.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
    L.agcgpu_comm_last_error.restype = C.c_char_p
    buf = (C.c_uint8 * 128)()
    if rank == 0 and L.agcgpu_comm_unique_id(buf) != 0:
        raise RuntimeError("agcgpu_comm_unique_id: " + L.agcgpu_comm_last_error().decode())
    t = torch.tensor(list(buf), dtype=torch.uint8, device=f"cuda:{device_index}" if dist.get_backend() == "nccl" else "cpu")
    dist.broadcast(t, src=0)
    ident = (C.c_uint8 * 128)(*t.cpu().tolist())
    if L.agcgpu_comm_init(rank, world, ident, device_index) != 0:
        raise RuntimeError("agcgpu_comm_init: " + L.agcgpu_comm_last_error().decode())
    return comm_stats()


class CommStats(C.Structure):
    _fields_ = [("nranks", C.c_uint32), ("rank", C.c_uint32), ("collectives", C.c_uint64), ("bytes_gathered", C.c_uint64)]


def comm_stats():
    from . import lib
    L = lib()
    st = CommStats()
    L.agcgpu_comm_get_stats.restype = C.c_int; L.agcgpu_comm_get_stats.argtypes = [C.POINTER(CommStats)]
    L.agcgpu_comm_get_stats(C.byref(st))
    return {"library": "NCCL (ncclAllGather from C++, device buffers)", "comm_nranks_seen": int(st.nranks), "rank": int(st.rank),
            "collectives": int(st.collectives), "bytes_gathered": int(st.bytes_gathered)}


def install_exchange_callback(L, device=None):
    """CPU tests (gloo): register torch.distributed's all-gather as the exchange step of library `L` through the callback entry
    (agcgpu_set_exchange).  The process group must be initialised.  device=None: CPU tensors.  Returns (rank, world)."""
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
