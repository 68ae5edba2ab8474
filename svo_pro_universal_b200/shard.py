"""Multi-GPU partitioning of independent work units (frame pairs, frames, features, seeds).

The hot path shards over independent units with NO collective inside it: rank r of W owns the contiguous block
[r*B/W, (r+1)*B/W). Only the per-unit results (a few hundred bytes each) are gathered to rank 0 at the end, with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def partition(n_units, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one unit and cover [0, n_units)."""
    lo = (n_units * rank) // world_size
    hi = (n_units * (rank + 1)) // world_size
    return lo, hi


def gather_to_rank0(local, n_units, group=None):
    """Gather equally-typed per-unit rows (numpy array [n_local, ...] or torch tensor) from all ranks onto rank 0 in global
    unit order. Returns the [n_units, ...] array on rank 0 and None elsewhere. Uses all_gather on padded blocks (blocks differ
    by at most one row), i.e. one small collective after the hot path."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    is_np = isinstance(local, np.ndarray)
    # bytes per row from the element type, so that a rank with 0 local rows (n_units < world size) still forms a [0, row_bytes] block
    if is_np:
        row_bytes = int(local.dtype.itemsize * np.prod(local.shape[1:], dtype=np.int64))
        t = torch.from_numpy(np.ascontiguousarray(local).reshape(-1).view(np.uint8).reshape(local.shape[0], row_bytes).copy())
    else:
        row_bytes = int(local.element_size() * np.prod(tuple(local.shape[1:]), dtype=np.int64))
        t = local.contiguous().reshape(-1).view(torch.uint8).reshape(local.shape[0], row_bytes)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = t.to(dev)
    max_rows = -(-n_units // world)
    pad = torch.zeros((max_rows, row_bytes), dtype=torch.uint8, device=dev)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != 0:
        return None
    rows = []
    for r in range(world):
        lo, hi = partition(n_units, world, r)
        rows.append(bufs[r][: hi - lo].cpu())
    full = torch.cat(rows, 0).numpy()
    if is_np:
        return full.view(local.dtype).reshape((n_units,) + local.shape[1:])
    return full
