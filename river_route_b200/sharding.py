"""
Basin sharding for multi-GPU runs.  Reaches only connect inside a drainage basin (one downstream per reach,
outlets have downstream_river_id < 0: river_route/tools.py:98-99), and the reference names watersheds as its
unit of parallelism (docs/references/parallelism.md:67-75).  Basins are bin-packed over the ranks by reach count
(LPT); a rank routes its own basins with its own plan and no collective in the time loop.
"""
from __future__ import annotations

import numpy as np

from .plan import label_basins


def shard_by_basin(down: np.ndarray, n_parts: int, part_id: int):
    """
    Returns (idx, local_down): the params-file indices of this rank's reaches in their original relative order
    (so every confluence keeps its upstreams in ascending order and fp64 sums are unchanged) and the downstream
    index vector re-expressed in local indices.
    """
    down = np.asarray(down, dtype=np.int32)
    if n_parts <= 1:
        return np.arange(down.shape[0]), down
    _, _, part = label_basins(down, n_parts)
    idx = np.flatnonzero(part == part_id)
    new_of_old = np.full(down.shape[0], -1, dtype=np.int64)
    new_of_old[idx] = np.arange(idx.shape[0])
    d = down[idx]
    local = np.where(d >= 0, new_of_old[np.where(d >= 0, d, 0)], -1).astype(np.int32)
    assert not np.any((d >= 0) & (local < 0)), 'a basin was cut across ranks'
    return idx, local


def bind_to_gpu_numa(device_index: int) -> bool:
    """
    Pin the calling process to the CPUs NVML reports as local to the GPU, so that pinned host buffers are first
    touched on the GPU's NUMA node (one process per GPU streams its own chunk of the arrays over its own PCIe
    link).  Best effort: returns False when NVML or the cpuset does not allow it.
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = int(vis.split(',')[device_index]) if vis else device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64 or 1)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
