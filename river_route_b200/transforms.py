"""
Array-level entry points of the two transforms that feed the router, both executed on the GPU:

* ``uh_convolve``       -- UnitHydrograph.convolve (river_route/uhkernels/UnitHydrograph.py:77-107)
* ``weights_transform`` -- the SpMM core + in-place tail of runoff_to_qlateral (river_route/runoff.py:292-337)
* ``Transform``         -- the same two, kept resident on the device in front of the router
                           (``Plan.runoff_route_host``: grid runoff -> discharge without the lateral inflows
                           leaving the GPU)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check


def uh_convolve(lateral: np.ndarray, kernel: np.ndarray, state: np.ndarray) -> np.ndarray:
    """
    (T, n) runoff depths convolved with the (n_ks, n) kernel; ``state`` (n_ks, n) carries the spill-over
    between calls and is updated in place exactly as UnitHydrograph.state is (:100-105).
    """
    lateral = np.ascontiguousarray(lateral, dtype=np.float64)
    kernel = np.ascontiguousarray(kernel, dtype=np.float64)
    if kernel.ndim != 2:
        raise ValueError('kernel must be a 2D array')
    if state.shape != kernel.shape:
        raise ValueError(f'state shape {state.shape} does not match kernel shape {kernel.shape}')
    if state.dtype != np.float64 or not state.flags.c_contiguous:
        raise TypeError('state must be a C-contiguous float64 array (it is updated in place)')
    if lateral.ndim != 2 or lateral.shape[1] != kernel.shape[1]:
        raise ValueError('lateral must have shape (t, n_basins)')
    T, n = lateral.shape
    out = np.empty((T, n), dtype=np.float64)
    check(lib.rr_uh_convolve_host(n, kernel.shape[0], T, _lib.as_f64p(lateral), n, _lib.as_f64p(kernel), n,
                                  _lib.as_f64p(state), n, _lib.as_f64p(out), n))
    return out


def weights_transform(indptr, indices, data, runoff_raw: np.ndarray, cumulative: bool = False,
                      force_positive: bool = False, area=None, keep_nan: bool = False) -> np.ndarray:
    """
    CSR weights (n_rivers x n_points, scipy layout: int32 indptr/indices, float64 data) applied to the gathered
    grid runoff (T, n_points), float32 or float64, followed by the reference's tail.  Returns (T, n_rivers) fp64.
    ``keep_nan`` skips the NaN -> 0 step (runoff.py:331-333) for callers that resample the series first (:316-329).
    """
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    if runoff_raw.dtype not in (np.float32, np.float64):
        runoff_raw = runoff_raw.astype(np.float64)
    x = np.ascontiguousarray(runoff_raw)
    T, n_points = x.shape
    n_rivers = indptr.shape[0] - 1
    if indices.size and int(indices.max()) >= n_points:
        raise ValueError('weight table refers to grid cells outside the gathered runoff array')
    a = None if area is None else np.ascontiguousarray(area, dtype=np.float64)
    y = np.empty((T, n_rivers), dtype=np.float64)
    check(lib.rr_weights_transform_host(n_rivers, n_points, T, _lib.as_i32p(indptr), _lib.as_i32p(indices),
                                        _lib.as_f64p(data), x.ctypes.data_as(C.c_void_p), int(x.dtype == np.float32),
                                        n_points, _lib.as_f64p(y), n_rivers, int(cumulative), int(bool(force_positive)) | (2 if keep_nan else 0),
                                        _lib.as_f64p(a) if a is not None else None))
    return y


class Transform:
    """
    Device-resident weight table (CSR over the plan's river order; scipy layout as ``build_weight_csr`` returns it)
    and, optionally, the unit-hydrograph kernel with its carry-over state (``rr_transform`` of include/rr_b200.h).
    """

    def __init__(self, indptr, indices, data, n_points: int, area=None, device: int = -1):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        if indptr.ndim != 1 or indptr.shape[0] < 2 or indices.shape != data.shape:
            raise ValueError('indptr / indices / data are not a CSR matrix')
        self.n_rivers = int(indptr.shape[0] - 1)
        self.n_points = int(n_points)
        self.n_ks = 0
        a = None if area is None else np.ascontiguousarray(area, dtype=np.float64)
        if a is not None and a.shape != (self.n_rivers,):
            raise ValueError('area must have one value per river segment')
        self._h = None
        handle = C.c_void_p()
        check(lib.rr_transform_create(self.n_rivers, self.n_points, _lib.as_i32p(indptr), _lib.as_i32p(indices),
                                      _lib.as_f64p(data), _lib.as_f64p(a) if a is not None else None, int(device),
                                      C.byref(handle)))
        self._h = handle

    def set_unit_hydrograph(self, kernel: np.ndarray, state: np.ndarray | None = None):
        kernel = np.ascontiguousarray(kernel, dtype=np.float64)
        if kernel.ndim != 2 or kernel.shape[1] != self.n_rivers:
            raise ValueError('kernel must have shape (n_kernel_steps, n_basins)')
        st = None
        if state is not None:
            st = np.ascontiguousarray(state, dtype=np.float64)
            if st.shape != kernel.shape:
                raise ValueError(f'state shape {st.shape} does not match kernel shape {kernel.shape}')
        check(lib.rr_transform_set_uh(self._h, kernel.shape[0], _lib.as_f64p(kernel), self.n_rivers,
                                      _lib.as_f64p(st) if st is not None else None, self.n_rivers))
        self.n_ks = int(kernel.shape[0])
        return self

    def uh_state(self) -> np.ndarray:
        """Carry-over state (n_kernel_steps, n_basins) as UnitHydrograph.state holds it."""
        st = np.empty((self.n_ks, self.n_rivers), dtype=np.float64)
        check(lib.rr_transform_get_uh_state(self._h, _lib.as_f64p(st), self.n_rivers))
        return st

    def close(self):
        if getattr(self, '_h', None):
            lib.rr_transform_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
