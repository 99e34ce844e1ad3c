"""Process-level runtime helpers (one process per GPU)."""
from __future__ import annotations

import os

from ._ffi import check, lib


def device_count() -> int:
    return lib().cb200_device_count()


def init(device: int | None = None) -> int:
    """Bind this process to one GPU (LOCAL_RANK by default)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(lib().cb200_init(device))
    return device


def shutdown() -> None:
    lib().cb200_shutdown()


def set_stream(cuda_stream_handle: int | None) -> None:
    check(lib().cb200_set_stream(cuda_stream_handle))


def synchronize() -> None:
    check(lib().cb200_synchronize())


def launch_count() -> int:
    return int(lib().cb200_launch_count())
