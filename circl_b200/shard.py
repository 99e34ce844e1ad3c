"""Index sharding of a batch across the GPUs of one box.

Every unit on this path (polynomial, encapsulation, signature) is independent, so rank r of G owns
the contiguous index range ``shard_range(n, r, G)`` and there is no exchange during compute.  The only
collective is the result gather to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests),
issued per chunk so that it overlaps the kernels of the next chunk.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of rank ``rank``; the first ``n % world`` ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total: int, dst: int = 0, group=None, async_op: bool = False, out=None):
    """Gather row-sharded results (``local`` = rows shard_range(n_total, rank, world)) to rank ``dst``
    in global index order.  Returns (tensor_or_None, work_handles)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    works = []
    if rank == dst:
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        lo = shard_range(n_total, dst, world)[0]
        out[lo:lo + sizes[dst]].copy_(local, non_blocking=True)
        for r in range(world):
            if r == dst or sizes[r] == 0:
                continue
            rlo = shard_range(n_total, r, world)[0]
            works.append(dist.irecv(out[rlo:rlo + sizes[r]], src=r, group=group))
    else:
        if sizes[rank]:
            works.append(dist.isend(local.contiguous(), dst=dst, group=group))
    if not async_op:
        for w in works:
            w.wait()
        works = []
    return (out if rank == dst else None), works
