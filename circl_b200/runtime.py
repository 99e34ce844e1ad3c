"""Process-level runtime helpers: one process per GPU (init) or one process driving several (init_devices)."""
from __future__ import annotations

import atexit
import os

from ._ffi import check, lib

_atexit_registered = False


def _shutdown_at_exit() -> None:
    global _atexit_registered
    if not _atexit_registered:  # worker threads and streams are released before the interpreter tears CUDA down
        atexit.register(shutdown)
        _atexit_registered = True


def device_count() -> int:
    return lib().cb200_device_count()


def init(device: int | None = None) -> int:
    """Bind this process to one GPU (LOCAL_RANK by default)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(lib().cb200_init(device))
    _shutdown_at_exit()
    return device


def init_devices(ndev: int = 0) -> int:
    """One process drives GPUs 0..ndev-1 (0: all visible); host-pointer batches are sharded by index inside the library."""
    check(lib().cb200_init_devices(ndev))
    _shutdown_at_exit()
    return int(lib().cb200_active_devices())


def active_devices() -> int:
    return int(lib().cb200_active_devices())


def bind_thread_to_device(device: int) -> int:
    """Pin the calling thread to the CPUs next to GPU `device` (NUMA-local pinned buffers for the host path)."""
    return int(lib().cb200_bind_thread_to_device(device))


def release_stream(cuda_stream_handle: int | None) -> None:
    check(lib().cb200_release_stream(cuda_stream_handle))


def host_batch(n: int, unit_bytes: int):
    """(n, unit_bytes) uint8 numpy array in pinned memory for a batch that host-pointer calls split over all active GPUs:
    each GPU's rows sit on that GPU's NUMA node (cb200_host_alloc_batch).  Release with host_free(array)."""
    import ctypes as C

    import numpy as np
    p = lib().cb200_host_alloc_batch(n, unit_bytes)
    if not p:
        raise MemoryError(lib().cb200_last_error().decode())
    a = np.frombuffer((C.c_uint8 * (n * unit_bytes)).from_address(p), dtype=np.uint8).reshape(n, unit_bytes)
    _host_batches[a.ctypes.data] = p
    return a


def host_free(a) -> None:
    p = _host_batches.pop(a.ctypes.data, None)
    if p is not None:
        lib().cb200_host_free(p)


_host_batches: dict = {}


def shutdown() -> None:
    lib().cb200_shutdown()


def set_stream(cuda_stream_handle: int | None) -> None:
    check(lib().cb200_set_stream(cuda_stream_handle))


def synchronize() -> None:
    check(lib().cb200_synchronize())


def launch_count() -> int:
    return int(lib().cb200_launch_count())
