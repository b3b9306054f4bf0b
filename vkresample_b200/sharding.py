"""Frame sharding across workers / GPUs -- the host-side logic of batched mode.

The reference spreads a folder of frames over ``numThreads`` worker threads that all drive the same
GPU (VkResample.cpp:1622-1629): worker ``t`` handles the 1-based files ``f*numThreads + t + 1`` for
``f = 0 .. numLocalFiles-1``.  Here the same striding maps whole frames to ranks (one process per GPU,
``bench.py``) or to CLI worker threads bound to devices (``-gpus N``).  Frames are independent, so there
is no collective on the data path; ``torch.distributed`` only carries the barrier and the
max-over-ranks of the timing.
"""
from __future__ import annotations

import math
from typing import List


def local_file_count(num_files: int, num_workers: int, worker: int) -> int:
    """numLocalFiles exactly as the reference computes it (VkResample.cpp:1622-1626)."""
    n = int(math.ceil(num_files / float(num_workers)))
    if (n - 1) * num_workers + worker > num_files - 1:
        n -= 1
    return n


def frames_for_worker(num_files: int, num_workers: int, worker: int) -> List[int]:
    """1-based file numbers handled by ``worker`` (file name ``%06d.png``, VkResample.cpp:1629)."""
    return [f * num_workers + worker + 1 for f in range(local_file_count(num_files, num_workers, worker))]


def device_for_worker(worker: int, num_devices: int, first_device: int = 0) -> int:
    """CLI ``-gpus N``: worker thread t drives device (d + t) mod N."""
    return (first_device + worker % num_devices) % max(num_devices, 1)


def aggregate_frames_per_s(frames_per_rank: int, world: int, max_seconds: float) -> float:
    """whole-job throughput: all ranks' frames over the slowest rank's time"""
    return world * frames_per_rank / max_seconds
