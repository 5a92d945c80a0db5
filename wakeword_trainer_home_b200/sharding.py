"""Clip-index sharding of a global batch across ranks (one process per GPU).

The feature path has no exchange step: every clip is independent (per-clip top_db maximum,
per-clip energies, per-clip masks), so ranks just take disjoint index ranges - the same
partition torch's DistributedSampler would give a DDP trainer - and no collective is issued.
"""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_seed(seed: int, rank: int, step: int = 0) -> int:
    """Distinct, reproducible RNG seed per (run seed, rank, step) for the augmentation draws;
    saving (seed, step) is all a resumed run needs to redraw the same augmentations."""
    return (seed * 1_000_003 + rank * 7919 + step * 104_729) % (2 ** 63 - 1)


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin the calling process to the CPU cores nearest to GPU `local_rank` (NVML's ideal affinity), so
    the pinned host buffers it allocates afterwards are NUMA-local to that GPU's PCIe root.  Matters
    only for the host-fed path when several ranks upload at once.  Returns False if NVML refuses
    (containers with a restricted cpuset)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        return True
    except Exception:
        return False


class numa_local:
    """Context manager: run the enclosed block (pinned-buffer allocation + first touch) on the CPU cores nearest
    to GPU ``local_rank``, then give the process its previous CPU affinity back - so that host-side work outside
    the block (e.g. a multi-threaded CPU baseline) still sees every core.  ``.bound`` tells whether NVML agreed."""

    def __init__(self, local_rank: int):
        self.local_rank, self.bound, self._prev = local_rank, False, None

    def __enter__(self):
        import os
        try:
            self._prev = os.sched_getaffinity(0)
        except Exception:
            self._prev = None
        self.bound = bind_to_gpu_numa(self.local_rank)
        return self

    def __exit__(self, *exc):
        import os
        if self.bound and self._prev:
            try:
                os.sched_setaffinity(0, self._prev)
            except Exception:
                pass
        return False
