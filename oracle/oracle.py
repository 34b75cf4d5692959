"""
CPU oracle for river-route's routing hot path  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``bench.py``'s cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module.  The product package
(``river_route_b200``) never does; it fails loudly without its CUDA library.

Parity status: PINNED against the reference's own numba / scipy code run in the
build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).

File:line citations are relative to ``/root/reference/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_i64 = C.c_int64
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)


def build() -> None:
    """Compile the C restatement (gcc, seconds)."""
    subprocess.run(['make', '-s', '-C', _HERE], check=True)


def _lib(fma: bool = False) -> C.CDLL:
    name = 'librr_oracle_fma.so' if fma else 'librr_oracle.so'
    if name not in _LIBS:
        path = os.path.join(_HERE, '_build', name)
        if not os.path.exists(path):
            build()
        _LIBS[name] = C.CDLL(path)
    return _LIBS[name]


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_pf64)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_pi32)


def _i64a(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_pi64)


def _rows(a: np.ndarray):
    """(pointer, leading dimension) of a C-contiguous-rows float64 2-D array."""
    assert a.dtype == np.float64 and a.ndim == 2 and (a.shape[1] <= 1 or a.strides[1] == 8)
    ld = a.strides[0] // 8 if a.shape[0] > 1 else max(a.shape[1], 1)
    return a.ctypes.data_as(_pf64), ld


# --------------------------------------------------------------------------------------
# Host-side derivations (exact restatements; numpy/scipy do the same arithmetic)
# --------------------------------------------------------------------------------------
def downstream_index(river_ids: np.ndarray, downstream_ids: np.ndarray) -> np.ndarray:
    """
    Index of each reach's downstream reach, -1 for outlets, with the reference's
    error behaviour.  river_route/tools.py:94-107: outlets are ``downstream_id < 0``;
    unknown ids raise ``ValueError('Unknown downstream_river_id: ...')``; a downstream
    index <= the upstream index raises the 'topologically sorted' ValueError.
    Errors are reported for the first offending row, as the reference's loop does.
    """
    river_ids = np.asarray(river_ids)
    downstream_ids = np.asarray(downstream_ids)
    lookup = {int(r): i for i, r in enumerate(river_ids.tolist())}
    down = np.full(river_ids.shape[0], -1, dtype=np.int64)
    for up, d in enumerate(downstream_ids.tolist()):
        if d < 0:
            continue
        if d not in lookup:
            raise ValueError(f'Unknown downstream_river_id: {d}')
        di = lookup[int(d)]
        if di <= up:
            raise ValueError('params_file must be topologically sorted upstream to downstream')
        down[up] = di
    return down


def csc_from_down(down: np.ndarray):
    """CSC (indptr, indices) of A[down[i], i] = 1  (tools.py:108-109; scipy int32)."""
    down = np.asarray(down, dtype=np.int64)
    has = down >= 0
    indptr = np.zeros(down.shape[0] + 1, dtype=np.int32)
    np.cumsum(has, out=indptr[1:])
    return indptr, down[has].astype(np.int32)


def muskingum_coefficients(k: np.ndarray, x: np.ndarray, dt_routing: float):
    """c1, c2, c3 with the reference's exact expressions (routers/Muskingum.py:174-185)."""
    dt_div_k = dt_routing / k
    denominator = dt_div_k + (2 * (1 - x))
    _2x = 2 * x
    c1 = (dt_div_k - _2x) / denominator
    c2 = (dt_div_k + _2x) / denominator
    c3 = ((2 * (1 - x)) - dt_div_k) / denominator
    if not np.allclose(c1 + c2 + c3, 1):
        raise ValueError('Muskingum coefficients do not sum to 1, check routing parameters and time step')
    return c1, c2, c3


def lhs_off_data(c1: np.ndarray, csc_indices: np.ndarray) -> np.ndarray:
    """Muskingum.py:192."""
    return np.ascontiguousarray(-c1[csc_indices])


def unit_split(down: np.ndarray):
    """
    Headwater / inner split and the two sub-adjacency CSCs of
    routers/UnitMuskingum.py:39-54, derived from the downstream-index vector.
    Returns dict with hw_idx, inner_idx (int64, ascending), A_inner CSC and
    A_hw_to_inner CSC (indptr, indices int32; data all 1.0).
    """
    n = down.shape[0]
    indeg = np.zeros(n, dtype=np.int64)
    np.add.at(indeg, down[down >= 0], 1)
    hw_mask = indeg == 0
    hw_idx = np.where(hw_mask)[0]
    inner_idx = np.where(~hw_mask)[0]
    pos_inner = np.full(n, -1, dtype=np.int64)
    pos_inner[inner_idx] = np.arange(inner_idx.shape[0])

    def sub_csc(cols):
        d = down[cols]
        has = d >= 0
        rows = np.where(has, pos_inner[np.where(has, d, 0)], -1)
        has &= rows >= 0
        indptr = np.zeros(cols.shape[0] + 1, dtype=np.int32)
        np.cumsum(has, out=indptr[1:])
        indices = rows[has].astype(np.int32)
        return indptr, indices, np.ones(indices.shape[0], dtype=np.float64)

    return dict(hw_idx=hw_idx, inner_idx=inner_idx, a_inner=sub_csc(inner_idx), a_hw=sub_csc(hw_idx))


# --------------------------------------------------------------------------------------
# Routing loops (C restatement)
# --------------------------------------------------------------------------------------
def muskingum_route(indptr, indices, lhs_off, c2, c3, q_t, discharge_array,
                    num_output_steps, num_routing_per_output, fma=False):
    """Same call shape as _numba_kernels.py:9-14; mutates q_t and discharge_array."""
    n = q_t.shape[0]
    indptr, p_indptr = _i32(indptr)
    indices, p_indices = _i32(indices)
    lhs_off, p_lhs = _f64(lhs_off)
    c2, p_c2 = _f64(c2)
    c3, p_c3 = _f64(c3)
    assert q_t.dtype == np.float64 and q_t.flags.c_contiguous
    p_out, ldo = _rows(discharge_array)
    rc = _lib(fma).rr_oracle_muskingum_route(
        _i64(n), p_indptr, p_indices, p_lhs, p_c2, p_c3, q_t.ctypes.data_as(_pf64),
        p_out, _i64(ldo), _i64(num_output_steps), _i64(num_routing_per_output))
    assert rc == 0


def rapid_route(indptr, indices, lhs_off, c2, c3, c4_dt, q_t, qlateral, discharge_array,
                num_substeps, fma=False):
    """Same call shape as _numba_kernels.py:50-55; mutates q_t and discharge_array."""
    n = q_t.shape[0]
    indptr, p_indptr = _i32(indptr)
    indices, p_indices = _i32(indices)
    lhs_off, p_lhs = _f64(lhs_off)
    c2, p_c2 = _f64(c2)
    c3, p_c3 = _f64(c3)
    c4_dt, p_c4 = _f64(c4_dt)
    qlateral = np.ascontiguousarray(qlateral, dtype=np.float64)
    assert q_t.dtype == np.float64 and q_t.flags.c_contiguous
    p_ql, ldq = _rows(qlateral)
    p_out, ldo = _rows(discharge_array)
    rc = _lib(fma).rr_oracle_rapid_route(
        _i64(n), p_indptr, p_indices, p_lhs, p_c2, p_c3, p_c4, q_t.ctypes.data_as(_pf64),
        p_ql, _i64(ldq), p_out, _i64(ldo), _i64(qlateral.shape[0]), _i64(num_substeps))
    assert rc == 0


def unit_route(lhs_indptr, lhs_indices, lhs_off,
               a_inner_indptr, a_inner_indices, a_inner_data,
               a_hw_indptr, a_hw_indices, a_hw_data,
               c1_inner, c2_inner, c3_inner, hw_idx, inner_idx,
               q_ch, q_full, convolved_lateral, discharge_array, num_substeps, fma=False):
    """Same call shape as _numba_kernels.py:89-99."""
    keep = []

    def k(pair):
        keep.append(pair[0])
        return pair[1]

    conv = np.ascontiguousarray(convolved_lateral, dtype=np.float64)
    p_conv, ldc = _rows(conv)
    p_out, ldo = _rows(discharge_array)
    assert q_ch.dtype == np.float64 and q_full.dtype == np.float64
    rc = _lib(fma).rr_oracle_unit_route(
        _i64(len(inner_idx)), _i64(len(hw_idx)),
        k(_i32(lhs_indptr)), k(_i32(lhs_indices)), k(_f64(lhs_off)),
        k(_i32(a_inner_indptr)), k(_i32(a_inner_indices)), k(_f64(a_inner_data)),
        k(_i32(a_hw_indptr)), k(_i32(a_hw_indices)), k(_f64(a_hw_data)),
        k(_f64(c1_inner)), k(_f64(c2_inner)), k(_f64(c3_inner)),
        k(_i64a(hw_idx)), k(_i64a(inner_idx)),
        q_ch.ctypes.data_as(_pf64), q_full.ctypes.data_as(_pf64),
        p_conv, _i64(ldc), p_out, _i64(ldo), _i64(conv.shape[0]), _i64(num_substeps))
    assert rc == 0


def uh_convolve(lateral, kernel, state, fma=False):
    """
    UnitHydrograph.convolve semantics (uhkernels/UnitHydrograph.py:77-107) in the
    summation order of convolve_incrementally (:64-75).  ``state`` (n_ks, n) is
    updated in place; returns the (T, n) convolved array.
    """
    lateral = np.ascontiguousarray(lateral, dtype=np.float64)
    kernel = np.ascontiguousarray(kernel, dtype=np.float64)
    assert state.dtype == np.float64 and state.flags.c_contiguous and state.shape == kernel.shape
    T, n = lateral.shape
    out = np.empty((T, n), dtype=np.float64)
    rc = _lib(fma).rr_oracle_uh_convolve(
        _i64(n), _i64(kernel.shape[0]), _i64(T),
        lateral.ctypes.data_as(_pf64), _i64(n), kernel.ctypes.data_as(_pf64), _i64(n),
        state.ctypes.data_as(_pf64), _i64(n), out.ctypes.data_as(_pf64), _i64(n))
    assert rc == 0
    return out


def weights_csr(river_idx, point_idx, values, n_rivers, n_points):
    """
    CSR of the weight matrix exactly as scipy builds it from COO triplets
    (runoff.py:292-295): duplicates summed, columns ascending within a row.
    Pure numpy so the oracle does not depend on scipy's version.
    """
    river_idx = np.asarray(river_idx, dtype=np.int64)
    point_idx = np.asarray(point_idx, dtype=np.int64)
    values = np.asarray(values, dtype=np.float64)
    order = np.lexsort((np.arange(river_idx.shape[0]), point_idx, river_idx))  # stable: input order kept in ties
    r, p, v = river_idx[order], point_idx[order], values[order]
    if r.shape[0]:
        first = np.ones(r.shape[0], dtype=bool)
        first[1:] = (r[1:] != r[:-1]) | (p[1:] != p[:-1])
    else:
        first = np.zeros(0, dtype=bool)
    starts = np.flatnonzero(first)
    # duplicates are added left to right in input order (scipy's coo->csr + sum_duplicates)
    data = np.add.reduceat(v, starts) if starts.shape[0] else np.zeros(0)
    rows, cols = r[first], p[first]
    indptr = np.zeros(n_rivers + 1, dtype=np.int32)
    np.add.at(indptr, rows + 1, 1)
    np.cumsum(indptr, out=indptr)
    assert cols.max(initial=-1) < n_points
    return indptr, cols.astype(np.int32), data


def weights_transform(indptr, indices, data, runoff_raw, cumulative=False, force_positive=False, area=None,
                      fma=False, keep_nan=False):
    """
    (T, n_points) gathered grid runoff -> (T, n_rivers) qlateral: the SpMM of
    runoff.py:298 plus the tail of :309-337 (cumulative diff, clip, NaN->0, x area).
    """
    x = np.ascontiguousarray(runoff_raw, dtype=np.float64)
    T = x.shape[0]
    n_rivers = len(indptr) - 1
    y = np.empty((T, n_rivers), dtype=np.float64)
    keep_i, p_i = _i32(indptr)
    keep_j, p_j = _i32(indices)
    keep_w, p_w = _f64(data)
    if area is not None:
        keep_a, p_a = _f64(area)
    else:
        p_a = None
    rc = _lib(fma).rr_oracle_weights_transform(
        _i64(n_rivers), _i64(T), p_i, p_j, p_w,
        x.ctypes.data_as(_pf64), _i64(x.shape[1]), y.ctypes.data_as(_pf64), _i64(n_rivers),
        C.c_int(int(cumulative)), C.c_int(int(bool(force_positive)) | (2 if keep_nan else 0)), p_a)
    assert rc == 0
    return y


# --------------------------------------------------------------------------------------
# The reference's recommended CPU parallelism: independent watersheds in separate
# workers (docs/references/parallelism.md:67-112).  Used only for the CPU baseline.
# --------------------------------------------------------------------------------------
def rapid_route_sharded(shards, num_substeps, threads):
    """
    shards: list of dicts with keys indptr, indices, lhs_off, c2, c3, c4_dt, q_t, ql, out.
    Each shard is an independent sub-network; ctypes releases the GIL so the C
    loops run concurrently on ``threads`` host cores.
    """
    def run(s):
        rapid_route(s['indptr'], s['indices'], s['lhs_off'], s['c2'], s['c3'], s['c4_dt'],
                    s['q_t'], s['ql'], s['out'], num_substeps)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(run, shards))
