"""
Deterministic synthetic inputs of the shapes named in BASELINE.json / SURVEY.md section 8d.
Used by tests and bench.py; not part of the routing path.
"""
from __future__ import annotations

import heapq

import numpy as np

from . import _lib
from ._lib import lib, check


def forest(n: int, n_basins: int = 1, seed: int = 0, depth_bias: float = 0.5, main_stem: int = 0,
           sigma: float = 1.5) -> np.ndarray:
    """
    Downstream-index vector (int32, -1 = outlet) of a forest grown upstream from its outlets; index =
    reverse growth order so ``down[i] > i`` (a valid params_file order).  Basins are contiguous index
    ranges with lognormal(sigma) sizes; ``main_stem`` pre-seeds a chain of that length in the first basin.
    """
    down = np.empty(n, dtype=np.int32)
    check(lib.rr_synth_forest(int(n), int(n_basins), int(seed), float(depth_bias), int(main_stem), float(sigma),
                              _lib.as_i32p(down)))
    return down


def ids_from_down(down: np.ndarray):
    """river_id = index + 1, downstream_river_id = -1 for outlets (the params_file columns)."""
    river_ids = np.arange(1, down.shape[0] + 1, dtype=np.int64)
    downstream = np.where(down >= 0, down.astype(np.int64) + 1, -1)
    return river_ids, downstream


def depth(down: np.ndarray) -> int:
    """Longest upstream-to-outlet path, in reaches."""
    lvl = levels(down)
    return int(lvl.max()) + 1 if lvl.size else 0


def levels(down: np.ndarray) -> np.ndarray:
    """Topological level of every reach (0 = headwater).  Vectorised over wavefronts of the forest."""
    n = down.shape[0]
    lvl = np.zeros(n, dtype=np.int32)
    indeg = np.bincount(down[down >= 0], minlength=n).astype(np.int32)
    frontier = np.flatnonzero(indeg == 0)
    while frontier.size:
        d = down[frontier]
        ok = d >= 0
        src, dst = frontier[ok], d[ok]
        np.maximum.at(lvl, dst, lvl[src] + 1)
        np.subtract.at(indeg, dst, 1)
        cand = np.unique(dst)
        frontier = cand[indeg[cand] == 0]
    return lvl


def relabel(down: np.ndarray, order: np.ndarray) -> np.ndarray:
    """Apply a new reach order: ``order[k]`` = old index of the reach placed at new index k."""
    n = down.shape[0]
    new_of_old = np.empty(n, dtype=np.int64)
    new_of_old[order] = np.arange(n)
    d_old = down[order]
    return np.where(d_old >= 0, new_of_old[np.where(d_old >= 0, d_old, 0)], -1).astype(np.int32)


def level_sorted_order(down: np.ndarray) -> np.ndarray:
    """Stable sort by topological level -- another valid params_file order (all headwaters first)."""
    return np.argsort(levels(down), kind='stable')


def random_topological_order(down: np.ndarray, seed: int = 0) -> np.ndarray:
    """A uniformly shuffled but valid upstream-before-downstream order (small networks only: Python loop)."""
    n = down.shape[0]
    rng = np.random.default_rng(seed)
    prio = rng.permutation(n)
    indeg = np.bincount(down[down >= 0], minlength=n)
    heap = [(int(prio[i]), int(i)) for i in np.flatnonzero(indeg == 0)]
    heapq.heapify(heap)
    order = np.empty(n, dtype=np.int64)
    k = 0
    while heap:
        _, i = heapq.heappop(heap)
        order[k] = i
        k += 1
        d = int(down[i])
        if d >= 0:
            indeg[d] -= 1
            if indeg[d] == 0:
                heapq.heappush(heap, (int(prio[d]), d))
    assert k == n
    return order


def muskingum_params(n: int, seed: int = 0):
    """k ~ U(1800, 20000) s, x ~ U(0.05, 0.4)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(1800.0, 20000.0, n), rng.uniform(0.05, 0.4, n)


def lateral_volumes(T: int, n: int, seed: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """gamma(0.3, 5e4 m3) with about half exact zeros, fp64, shape (T, n)."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty((T, n), dtype=np.float64)
    step = max(1, (1 << 24) // max(n, 1))
    for t0 in range(0, T, step):
        t1 = min(T, t0 + step)
        blk = rng.gamma(0.3, 5.0e4, size=(t1 - t0, n))
        blk[rng.random((t1 - t0, n)) < 0.5] = 0.0
        out[t0:t1, :n] = blk
    return out
