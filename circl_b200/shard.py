"""Index sharding of a batch across the GPUs of one box.

Every unit on this path (polynomial, encapsulation, signature) is independent, so rank r of G owns
the contiguous index range ``shard_range(n, r, G)`` and there is no exchange during compute.  The only
exchange is the result gather to rank 0, issued per chunk so that it overlaps the kernels of the next chunk
(``RowGather``: peer copies by the copy engines into rank 0's CUDA-IPC-mapped buffer on GPUs, plain
send/recv under gloo in the CPU tests -- same layout code, other transport).

One process driving several GPUs needs none of this: ``cb200_init_devices`` shards host-pointer batches inside
the library (circl_b200/csrc/api.cu, for_each_shard).
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


class _RawCuda:
    """A raw device pointer as a __cuda_array_interface__ object (torch.as_tensor views it without a copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class RowGather:
    """Weak-scaling gather for the one-process-per-GPU launch: every rank produces ``n`` rows of each of several
    byte matrices (row widths ``widths``); rank 0 ends up with ``world * n`` rows of each in global index order
    (rank r's row i is global row r*n + i).

    transport "ipc" (GPU): rank 0 allocates the destination with cb200_gather_alloc and broadcasts its CUDA IPC handle
    through torch.distributed; ``push`` is cb200_gather_push -- a cudaMemcpyAsync into the mapped peer memory on a copy
    stream that waits for the kernels queued so far: NVLink traffic by the copy engines, no SM, no NCCL kernel.
    transport "sendrecv" (CPU tests under gloo): the same row placement over dist.send / dist.recv.
    """

    def __init__(self, n: int, widths, transport: str = "ipc", group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        self.n, self.widths, self.transport, self.group = n, list(widths), transport, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.offsets, total = [], 0
        for w in self.widths:
            self.offsets.append(total)
            total += self.world * n * w
        self.total = total
        self.owner = self.rank == 0
        self._pending = []
        if transport == "ipc":
            from ._ffi import check, lib
            self._L, self._check = lib(), check
            handle = (ctypes.c_uint8 * 64)()
            self.base = ctypes.c_void_p()
            if self.owner:
                check(self._L.cb200_gather_alloc(total, ctypes.byref(self.base), handle))
            box = [bytes(handle)]
            dist.broadcast_object_list(box, src=0, group=group)
            if not self.owner:
                h = (ctypes.c_uint8 * 64).from_buffer_copy(box[0])
                check(self._L.cb200_gather_open(h, ctypes.byref(self.base)))
            self.view = torch.as_tensor(_RawCuda(self.base.value, total), device="cuda") if self.owner else None
        elif transport == "sendrecv":
            self.view = torch.zeros(total, dtype=torch.uint8) if self.owner else None
        else:
            raise ValueError(transport)

    def _dst_offset(self, which: int, rank: int, lo: int) -> int:
        return self.offsets[which] + (rank * self.n + lo) * self.widths[which]

    def push(self, tensors, lo: int, hi: int):
        """rows [lo, hi) of this rank's matrices -> their place in the global order on rank 0"""
        import torch.distributed as dist
        for which, t in enumerate(tensors):
            w = self.widths[which]
            if self.transport == "ipc":
                self._check(self._L.cb200_gather_push(self.base.value + self._dst_offset(which, self.rank, lo),
                                                      t[lo:hi].data_ptr(), (hi - lo) * w))
            elif self.owner:
                o = self._dst_offset(which, 0, lo)
                self.view[o:o + (hi - lo) * w] = t[lo:hi].reshape(-1)
                for r in range(1, self.world):
                    o = self._dst_offset(which, r, lo)
                    self._pending.append(dist.irecv(self.view[o:o + (hi - lo) * w], src=r, group=self.group))
            else:
                self._pending.append(dist.isend(t[lo:hi].reshape(-1).contiguous(), dst=0, group=self.group))

    def dst_ptr(self, which: int, lo: int = 0) -> int:
        """ipc: raw address of this rank's row `lo` of matrix `which` in rank 0's buffer (for cb200_*_push entry points)"""
        return self.base.value + self._dst_offset(which, self.rank, lo)

    def flush(self, any_tensor=None, on_host: bool = False):
        """ipc: make the caller's stream (or the host) wait for this rank's pushes; sendrecv: wait for the requests"""
        if self.transport == "ipc":
            self._check(self._L.cb200_gather_flush(any_tensor.data_ptr(), 1 if on_host else 0))
        else:
            for w in self._pending:
                w.wait()
            self._pending = []

    def matrix(self, which: int):
        """rank 0: the gathered (world*n, width) matrix `which` (a view of the destination buffer)"""
        w, off = self.widths[which], self.offsets[which]
        return self.view[off:off + self.world * self.n * w].view(self.world * self.n, w)

    def close(self):
        if self.transport != "ipc":
            return
        if self.owner:
            self.view = None
            self._L.cb200_gather_free(self.base)
        else:
            self._L.cb200_gather_close(self.base)
