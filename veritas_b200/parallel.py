"""Host-side plumbing of the multi-GPU x-slab decomposition: one process per GPU, equal contiguous
x-slabs of the single logical patch, NCCL communicator bootstrap through whatever launcher started the ranks
(torch.distributed here: bench.py, tests).  No data-path logic lives here — halo exchange and the moment all-gather are
issued by the library itself (csrc/vrt_comm.cu)."""
import ctypes as C


def slab_bounds(nx, rank, world):
    """Columns [x_begin, x_end) owned by `rank`: equal slabs, ordered by rank (vrt_set_slab requires exactly this)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if nx % world:
        raise ValueError(f"x size {nx} is not divisible by {world} ranks")
    n = nx // world
    if n < 8:
        raise ValueError("slabs must hold at least 8 columns")
    return rank * n, (rank + 1) * n


def broadcast_unique_id(dist, lib, rank, device=None):
    """Rank 0 asks the library for an ncclUniqueId (128 bytes); everyone receives it through the launcher's process
    group (`dist` = torch.distributed, any backend).  Returns the 128 bytes."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        if lib.vrt_nccl_unique_id(raw) != 0:
            raise RuntimeError("vrt_nccl_unique_id failed (libnccl not loadable)")
        buf = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def max_over_ranks(dist, value, device=None):
    """Timing convention of the bench: the slowest rank defines the step time."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
