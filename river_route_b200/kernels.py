"""
Drop-in replacements for the three numba kernels of river_route/routers/_numba_kernels.py.

Same names, same positional arguments, same in-place mutation of ``q_t`` / ``discharge_array`` and
no return value, so a maintainer can switch the reference over with three import lines
(INTEGRATION.md).  The arrays are marshalled to librr_b200.so; nothing is computed on the CPU.
"""
from __future__ import annotations

import numpy as np

from .plan import MODE_MUSKINGUM, MODE_RAPID, MODE_UNIT, Plan, down_from_csc

__all__ = ['muskingum_route', 'rapid_route', 'unit_route', 'clear_plan_cache']

_PLANS: dict = {}
_MAX_PLANS = 4


def clear_plan_cache() -> None:
    for p in _PLANS.values():
        p.close()
    _PLANS.clear()


def _plan_for(down: np.ndarray) -> Plan:
    """The reference re-uses its CSC arrays across files; the plan is cached on the same key."""
    key = (down.shape[0], hash(down.tobytes()))
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= _MAX_PLANS:
            _PLANS.pop(next(iter(_PLANS))).close()
        plan = _PLANS[key] = Plan(down)
    return plan


def _c1_from_lhs(n: int, indices, lhs_off_data) -> np.ndarray:
    """lhs_off_data[j] = -c1[row(j)] (routers/Muskingum.py:192); rows without upstreams never use c1."""
    c1 = np.zeros(n, dtype=np.float64)
    c1[np.asarray(indices)] = -np.asarray(lhs_off_data, dtype=np.float64)
    return c1


def _check_state(q_t, n_name='q_t'):
    if not isinstance(q_t, np.ndarray) or q_t.dtype != np.float64 or q_t.ndim != 1 or not q_t.flags.c_contiguous:
        raise TypeError(f'{n_name} must be a contiguous 1-D float64 numpy array (it is updated in place)')


def muskingum_route(csc_indptr, csc_indices, lhs_off_data, c2, c3, q_t, discharge_array,
                    num_output_steps, num_routing_per_output) -> None:
    """Muskingum channel-only loop -- replaces _numba_kernels.py:9-46."""
    _check_state(q_t)
    n = q_t.shape[0]
    if discharge_array.shape != (num_output_steps, n):
        raise ValueError('discharge_array must have shape (num_output_steps, n)')
    plan = _plan_for(down_from_csc(csc_indptr, csc_indices, n))
    plan.set_coefficients(_c1_from_lhs(n, csc_indices, lhs_off_data), c2, c3, None)
    plan.route_host(MODE_MUSKINGUM, q_t, None, discharge_array, int(num_routing_per_output))


def rapid_route(csc_indptr, csc_indices, lhs_off_data, c2, c3, c4_dt, q_t, qlateral, discharge_array,
                num_substeps) -> None:
    """RapidMuskingum loop with lateral inflow -- replaces _numba_kernels.py:49-84."""
    _check_state(q_t)
    n = q_t.shape[0]
    if qlateral.ndim != 2 or qlateral.shape[1] != n:
        # the numba kernel would silently read out of bounds here (no bounds checks)
        raise ValueError('qlateral must have shape (num_runoff_steps, n)')
    if discharge_array.shape != qlateral.shape:
        raise ValueError('discharge_array must have the same shape as qlateral')
    plan = _plan_for(down_from_csc(csc_indptr, csc_indices, n))
    plan.set_coefficients(_c1_from_lhs(n, csc_indices, lhs_off_data), c2, c3, c4_dt)
    plan.route_host(MODE_RAPID, q_t, qlateral, discharge_array, int(num_substeps))


def unit_route(lhs_indptr, lhs_indices, lhs_off_data,
               a_inner_indptr, a_inner_indices, a_inner_data,
               a_hw_indptr, a_hw_indices, a_hw_data,
               c1_inner, c2_inner, c3_inner,
               hw_idx, inner_idx,
               q_ch, q_full,
               convolved_lateral, discharge_array, num_substeps) -> None:
    """
    UnitMuskingum loop -- replaces _numba_kernels.py:88-171.  ``q_ch`` and ``q_full`` are the
    inner-reach vectors of the reference; both are updated in place.
    """
    _check_state(q_ch, 'q_ch')
    _check_state(q_full, 'q_full')
    hw_idx = np.asarray(hw_idx, dtype=np.int64)
    inner_idx = np.asarray(inner_idx, dtype=np.int64)
    n = discharge_array.shape[1]
    if hw_idx.shape[0] + inner_idx.shape[0] != n:
        raise ValueError('hw_idx and inner_idx must partition the river segments')
    if np.any(np.asarray(a_inner_data) != 1.0) or np.any(np.asarray(a_hw_data) != 1.0):
        raise ValueError('adjacency data must be 1.0 (river_route/tools.py:108)')
    # global downstream index from the two sub-adjacency matrices (UnitMuskingum.py:45-46)
    down = np.full(n, -1, dtype=np.int32)
    d_in = down_from_csc(a_inner_indptr, a_inner_indices, inner_idx.shape[0])
    d_hw = down_from_csc(a_hw_indptr, a_hw_indices, hw_idx.shape[0])
    down[inner_idx[d_in >= 0]] = inner_idx[d_in[d_in >= 0]]
    down[hw_idx[d_hw >= 0]] = inner_idx[d_hw[d_hw >= 0]]
    if not np.array_equal(np.asarray(lhs_indptr), np.asarray(a_inner_indptr)) or \
            not np.array_equal(np.asarray(lhs_indices), np.asarray(a_inner_indices)):
        raise ValueError('LHS sparsity must equal A_inner (UnitMuskingum.py:68-70)')

    def full(v):
        a = np.zeros(n, dtype=np.float64)
        a[inner_idx] = v
        return a

    c1 = full(c1_inner)
    if lhs_off_data is not None and len(lhs_off_data):
        if not np.array_equal(np.asarray(lhs_off_data), -np.asarray(c1_inner)[np.asarray(lhs_indices)]):
            raise ValueError('lhs_off_data must equal -c1_inner[indices] (UnitMuskingum.py:70)')
    plan = _plan_for(down)
    plan.set_coefficients(c1, full(c2_inner), full(c3_inner), None)
    qc, qf = full(q_ch), full(q_full)
    plan.route_host(MODE_UNIT, qc, convolved_lateral, discharge_array, int(num_substeps), q_full=qf)
    q_ch[:] = qc[inner_idx]
    q_full[:] = qf[inner_idx]
