"""Frame sharding for batches of independent frames (BASELINE config 4).

The path does not shard inside a scan, so multi-GPU = partition the independent frames over the
ranks (one process per GPU), no collective in the data path, and ONE gather of the 32-byte
result records (pose[6] f32, iters i32, flags u32) to rank 0 at the end.
"""
import numpy as np

RESULT_DTYPE = np.dtype([("pose", "<f4", (6,)), ("iters", "<i4"), ("flags", "<u4")])


def frame_range(rank, world, total):
    """Contiguous block of frames for `rank`: frame f -> rank floor(f * world / total)."""
    lo = (rank * total + world - 1) // world
    hi = ((rank + 1) * total + world - 1) // world
    return lo, hi


def pack_results(res):
    """structured results -> float32 [n, 8] (bit-preserving) for a collective."""
    return np.ascontiguousarray(res).view(np.float32).reshape(-1, 8)


def unpack_results(buf):
    return np.ascontiguousarray(buf, dtype=np.float32).reshape(-1, 8).view(RESULT_DTYPE).reshape(-1)


def gather_results(local, dist=None, device=None):
    """Gather every rank's result records on rank 0 (returns None elsewhere).  `local` may be a numpy
    structured array (CPU / gloo) or a float32 torch tensor [n, 8] already on the rank's GPU (nccl)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local if isinstance(local, np.ndarray) else unpack_results(local.cpu().numpy())
    world, rank = dist.get_world_size(), dist.get_rank()
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(pack_results(local).copy())
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    pad = torch.zeros((m, 8), dtype=torch.float32, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return unpack_results(torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)]).cpu().numpy())
