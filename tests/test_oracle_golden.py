"""
Pins the CPU oracle (oracle/) against outputs of the reference's own code (tests/golden/, produced by
oracle/make_golden.py from /root/reference).  Everything here runs on the CPU.

Tolerances: the oracle keeps the reference's summation order, so differences come only from numba's
fastmath FMA contraction (SURVEY.md 7 "fastmath/FMA"); 1e-12 in the parity measure of SURVEY.md 8d is
~100x looser than observed and 100x tighter than the 1e-10 the GPU path is held to.
"""
import numpy as np
import pytest

from oracle import oracle
from tests.conftest import load_golden
from tests.helpers import parity_error

TOL = 1e-12


def test_topology_and_coefficients_bit_exact(route_golden):
    g = route_golden
    ids = g['river_ids']
    down = g['down']
    downstream_ids = np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1)
    assert np.array_equal(oracle.downstream_index(ids, downstream_ids), down)
    indptr, indices = oracle.csc_from_down(down)
    assert indptr.dtype == g['csc_indptr'].dtype == np.int32
    assert np.array_equal(indptr, g['csc_indptr']) and np.array_equal(indices, g['csc_indices'])
    c1, c2, c3 = oracle.muskingum_coefficients(g['k'], g['x'], int(g['dt_routing']))
    for mine, ref in ((c1, g['c1']), (c2, g['c2']), (c3, g['c3'])):
        assert np.array_equal(mine, ref)  # same numpy expressions -> identical bits
    assert np.array_equal(oracle.lhs_off_data(c1, indices), g['lhs_off'])
    assert np.array_equal((c1 + c2) / int(g['dt_runoff']), g['c4_dt'])
    sp = oracle.unit_split(down)
    assert np.array_equal(sp['hw_idx'], g['hw_idx']) and np.array_equal(sp['inner_idx'], g['inner_idx'])
    assert np.array_equal(sp['a_inner'][0], g['a_inner_indptr']) and np.array_equal(sp['a_inner'][1], g['a_inner_indices'])
    assert np.array_equal(sp['a_hw'][0], g['a_hw_indptr']) and np.array_equal(sp['a_hw'][1], g['a_hw_indices'])


@pytest.mark.parametrize('fma', [False, True])
def test_rapid_route_matches_reference(route_golden, fma):
    g = route_golden
    q = g['q0'].copy()
    out = np.zeros_like(g['rapid_out'])
    oracle.rapid_route(g['csc_indptr'], g['csc_indices'], g['lhs_off'], g['c2'], g['c3'], g['c4_dt'], q, g['ql'], out,
                       int(g['substeps']), fma=fma)
    assert parity_error(out, g['rapid_out']) < TOL
    assert parity_error(q, g['rapid_q']) < TOL
    assert np.array_equal(out == 0, g['rapid_out'] == 0)  # identical clamp pattern


@pytest.mark.parametrize('fma', [False, True])
def test_muskingum_route_matches_reference(route_golden, fma):
    g = route_golden
    q = g['q0'].copy()
    out = np.zeros_like(g['musk_out'])
    oracle.muskingum_route(g['csc_indptr'], g['csc_indices'], g['lhs_off'], g['c2'], g['c3'], q, out,
                           int(g['musk_nout']), int(g['musk_nrpo']), fma=fma)
    assert parity_error(out, g['musk_out']) < TOL
    assert parity_error(q, g['musk_q']) < TOL


def _unit_args(g):
    inner = g['inner_idx']
    c1i, c2i, c3i = g['c1'][inner], g['c2'][inner], g['c3'][inner]
    ones_i = np.ones(g['a_inner_indices'].shape[0])
    ones_h = np.ones(g['a_hw_indices'].shape[0])
    lhs = np.ascontiguousarray(-c1i[g['a_inner_indices']])  # UnitMuskingum.py:70
    return (g['a_inner_indptr'], g['a_inner_indices'], lhs, g['a_inner_indptr'], g['a_inner_indices'], ones_i,
            g['a_hw_indptr'], g['a_hw_indices'], ones_h, c1i, c2i, c3i, g['hw_idx'], inner)


@pytest.mark.parametrize('fma', [False, True])
def test_unit_route_matches_reference(route_golden, fma):
    g = route_golden
    inner, hw = g['inner_idx'], g['hw_idx']
    q_ch = g['q0'][inner].copy()
    q_full = q_ch.copy()
    out = np.zeros_like(g['unit_out'])
    oracle.unit_route(*_unit_args(g), q_ch, q_full, g['conv'], out, int(g['substeps']), fma=fma)
    q_final = np.empty_like(g['q0'])
    q_final[hw] = g['conv'][-1][hw]
    q_final[inner] = q_full
    assert parity_error(out, g['unit_out']) < TOL
    assert parity_error(q_final, g['unit_q']) < TOL


def test_uh_convolve_inside_router_matches_reference(route_golden):
    """UnitHydrograph.convolve as called from UnitMuskingum._router, two files in sequence (state chains)."""
    g = route_golden
    state = g['uh_state0'].copy()
    conv = oracle.uh_convolve(g['depths'], g['uh_kernel'], state)
    # fftconvolve noise is ~1e-16 of the column scale and large relative to exact zeros: normwise check
    col = np.max(np.abs(g['conv']), axis=0) + np.max(np.abs(g['uh_state0']), axis=0)
    assert parity_error(conv, g['conv'], col) < TOL
    assert parity_error(state, g['uh_state1'], col) < TOL
    oracle.uh_convolve(g['depths2'], g['uh_kernel'], state)
    assert parity_error(state, g['uh_state2'], col) < TOL


def test_uh_reference_known_answers():
    g = load_golden('uh.npz')
    # tests/test_uhkernels.py:52-78 (seed 123): full == incremental to 1e-12; the oracle follows the incremental order
    state = np.zeros_like(g['kernel'])
    conv = oracle.uh_convolve(g['lateral'], g['kernel'], state)
    assert np.array_equal(conv, g['conv_inc'])          # same operation order -> same bits as convolve_incrementally
    np.testing.assert_allclose(conv, g['conv_full'], rtol=1e-12)
    assert np.array_equal(state, g['state_inc'])
    # tests/test_uhkernels.py:81-99: impulse reproduces the kernel
    ker = g['impulse_kernel']
    lat = np.zeros((5, 2))
    lat[0, :] = 1.0
    res = oracle.uh_convolve(lat, ker, np.zeros_like(ker))
    np.testing.assert_allclose(res[:3], ker, rtol=1e-12)
    np.testing.assert_allclose(res[3:], 0.0, atol=1e-15)
    # carry-over across three calls including T < n_ks and T == 1
    state = np.zeros_like(g['kernel_b'])
    for c in range(3):
        out = oracle.uh_convolve(g[f'call{c}'], g['kernel_b'], state)
        assert np.array_equal(out, g[f'inc{c}'])
        col = np.max(np.abs(g['out0']), axis=0)
        assert parity_error(out, g[f'out{c}'], col) < TOL
        assert parity_error(state, g[f'state{c}'], col) < TOL
    assert np.array_equal(state, g['state_inc_final'])


@pytest.mark.parametrize('unit', ['m', 'mm'])
def test_weight_csr_and_transform_match_scipy(unit):
    g = load_golden('weights.npz')
    factor = 1 if unit == 'm' else .001
    n_riv, n_pts = len(g['river_ids_ordered']), g['runoff_raw'].shape[1]
    indptr, indices, data = oracle.weights_csr(g['river_idx'], g['point_idx'], g['proportion'] * factor, n_riv, n_pts)
    assert np.array_equal(indptr, g[f'csr_indptr_{unit}']) and np.array_equal(indices, g[f'csr_indices_{unit}'])
    assert np.array_equal(data, g[f'csr_data_{unit}'])
    for cumulative in (False, True):
        src = g['runoff_raw_cumulative'] if cumulative else g['runoff_raw']
        for vol in (False, True):
            ql = oracle.weights_transform(indptr, indices, data, src, cumulative=cumulative,
                                          area=g['catchment_area'] if vol else None)
            ref = g[f'ql_{unit}_cum{int(cumulative)}_vol{int(vol)}']
            assert not np.isnan(ql).any()
            assert parity_error(ql, ref) < TOL


def test_irregular_time_axis_resample_matches_reference():
    """runoff.py:316-337 with NaN cells: the NaNs run through cumsum / resample / interpolate and are zeroed after."""
    from river_route_b200.runoff import resample_irregular   # host-side pandas tail of the product (no GPU work)
    g, gi = load_golden('weights.npz'), load_golden('weights_irregular.npz')
    indptr, indices, data = g['csr_indptr_m'], g['csr_indices_m'], g['csr_data_m']
    t_in = gi['time_index'].astype('datetime64[s]')
    for cumulative in (False, True):
        src = gi['runoff_raw_cumulative'] if cumulative else gi['runoff_raw']
        assert np.isnan(src).any()
        for vol in (False, True):
            ql = oracle.weights_transform(indptr, indices, data, src, cumulative=cumulative, keep_nan=True)
            assert np.isnan(ql).any()
            out, t_out = resample_irregular(ql, t_in, g['river_ids_ordered'], g['catchment_area'] if vol else None)
            ref = gi[f'ql_cum{int(cumulative)}_vol{int(vol)}']
            assert out.shape == ref.shape and np.array_equal(t_out.astype('datetime64[s]').astype(np.int64), gi['time_out'])
            assert parity_error(out, ref) < TOL
            assert np.array_equal(out == 0.0, ref == 0.0)


def test_tools_known_answers():
    g = load_golden('tools.npz')
    down = oracle.downstream_index(g['ids9'], g['ds9'])  # 9-reach network of docs/references/math.md:70-80
    indptr, indices = oracle.csc_from_down(down)
    assert np.array_equal(indptr, g['A9_indptr']) and np.array_equal(indices, g['A9_indices'])
    dense = np.zeros((9, 9))
    dense[down[down >= 0], np.flatnonzero(down >= 0)] = 1.0
    assert np.array_equal(dense, g['A9_dense'])
    with pytest.raises(ValueError, match='topologically sorted'):          # tests/test_tools.py:48-53
        oracle.downstream_index(np.array([10, 20, 30]), np.array([20, -1, 10]))
    with pytest.raises(ValueError, match='Unknown downstream_river_id'):   # tests/test_tools.py:56-60
        oracle.downstream_index(np.array([10, 20]), np.array([-1, 999]))
    assert str(g['err_unsorted']) == 'params_file must be topologically sorted upstream to downstream'
    assert str(g['err_unknown']) == 'Unknown downstream_river_id: 999'
