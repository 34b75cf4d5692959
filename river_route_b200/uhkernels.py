"""
UnitHydrograph with the interface of river_route/uhkernels/UnitHydrograph.py; the convolution runs on
the GPU (rr_uh_convolve_* of include/rr_b200.h).  File I/O (scipy-sparse npz kernel, parquet state) is
host Python as in the reference.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .transforms import uh_convolve

__all__ = ['UnitHydrograph']


class UnitHydrograph:
    """
    Stateful runoff transformer driven by a precomputed (n_kernel_steps, n_basins) kernel.  ``convolve`` handles
    a whole (t, n_basins) series, ``convolve_incrementally`` one time step; both carry the spill-over between
    calls in ``state`` exactly as the reference does (:64-107).
    """

    def __init__(self, kernel_file=None, kernel: np.ndarray | None = None):
        if kernel is None:
            import scipy.sparse
            kernel = scipy.sparse.load_npz(kernel_file).toarray()
        self.kernel = np.ascontiguousarray(kernel, dtype=np.float64)
        if self.kernel.ndim != 2:
            raise ValueError('kernel must be a 2D array')
        self.reset_state()

    def reset_state(self) -> None:
        self.state = np.zeros_like(self.kernel, dtype=np.float64)

    def set_state(self, path):
        """State parquet has shape (n_basins, n_kernel_steps) -- basins as rows (:47-57)."""
        state = pd.read_parquet(path).T.to_numpy(dtype=np.float64, copy=True)
        if state.shape != self.kernel.shape:
            raise ValueError(f'state shape {state.shape} does not match kernel shape {self.kernel.shape}')
        self.state = np.ascontiguousarray(state)
        return self

    def write_state(self, path) -> None:
        pd.DataFrame(self.state.T).to_parquet(path)

    def convolve(self, lateral: np.ndarray) -> np.ndarray:
        """(t, n_basins) runoff depths -> (t, n_basins) convolved lateral inflow; updates ``state`` (:77-107)."""
        if not self.state.flags.c_contiguous or self.state.dtype != np.float64:
            self.state = np.ascontiguousarray(self.state, dtype=np.float64)
        return uh_convolve(lateral, self.kernel, self.state)

    def convolve_incrementally(self, runoff_vector: np.ndarray) -> np.ndarray:
        """One time step (:64-75); identical to ``convolve`` on a one-row series."""
        return self.convolve(np.asarray(runoff_vector, dtype=np.float64)[np.newaxis, :])[0]
