"""Multi-GPU plumbing for the query path: one process per GPU, torch.distributed only as the
transport.  Queries shard with no data-path collective; the partitioned-surface
closest-point case needs exactly one exchange, an elementwise MIN over the ranks' partial
squared distances (the reference's MPI ring of Conduit nodes,
quest/detail/DistributedClosestPointImpl.hpp:737-851, collapses to this on NVSwitch)."""
import numpy as np


def slab_range(n, rank, world):
    """contiguous [k0,k1) share of n slabs/items for `rank` of `world` (differs by at most one)"""
    return (n * rank) // world, (n * (rank + 1)) // world


def morton_partition(centroids, parts):
    """split surface cells into `parts` spatially coherent groups: contiguous ranges of the cells'
    Morton order (10 bits/dim over the centroid bounds).  Returns a list of index arrays."""
    c = np.asarray(centroids, np.float64)
    lo, hi = c.min(axis=0), c.max(axis=0)
    ext = np.where(hi - lo > 0, hi - lo, 1.0)
    q = np.minimum((1023 * (c - lo) / ext).astype(np.uint64), 1023)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x030000FF)
        v = (v | (v << 8)) & np.uint64(0x0300F00F)
        v = (v | (v << 4)) & np.uint64(0x030C30C3)
        v = (v | (v << 2)) & np.uint64(0x09249249)
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    order = np.argsort(code, kind="stable")
    return [order[slab_range(len(order), r, parts)[0]:slab_range(len(order), r, parts)[1]] for r in range(parts)]


def allreduce_min_(t):
    """in-place elementwise MIN over ranks (NCCL on device tensors, gloo on CPU tensors)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t


def allreduce_max_scalar(v, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
