"""GPU parity of the unit-hydrograph convolution and the grid-weight transform (through the C ABI)."""
import numpy as np
import pytest

from river_route_b200.transforms import uh_convolve, weights_transform
from oracle import oracle
from tests.conftest import load_golden, require_cuda
from tests.helpers import parity_error

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True)
def _cuda():
    require_cuda()


def test_uh_golden_known_answers():
    g = load_golden('uh.npz')
    state = np.zeros_like(g['kernel'])
    conv = uh_convolve(g['lateral'], g['kernel'], state)
    np.testing.assert_allclose(conv, g['conv_full'], rtol=1e-12)      # tests/test_uhkernels.py:52-78
    np.testing.assert_allclose(conv, g['conv_inc'], rtol=1e-12)
    np.testing.assert_allclose(state, g['state_inc'], rtol=1e-12, atol=1e-300)
    ker = g['impulse_kernel']                                         # tests/test_uhkernels.py:81-99
    lat = np.zeros((5, 2))
    lat[0, :] = 1.0
    res = uh_convolve(lat, ker, np.zeros_like(ker))
    np.testing.assert_allclose(res[:3], ker, rtol=1e-12)
    np.testing.assert_allclose(res[3:], 0.0, atol=1e-15)
    state = np.zeros_like(g['kernel_b'])                              # carry-over, T < n_ks, T == 1
    col = np.max(np.abs(g['out0']), axis=0)
    for c in range(3):
        out = uh_convolve(g[f'call{c}'], g['kernel_b'], state)
        assert parity_error(out, g[f'inc{c}'], col) < TOL
        assert parity_error(out, g[f'out{c}'], col) < TOL
        assert parity_error(state, g[f'state{c}'], col) < TOL


@pytest.mark.parametrize('n,n_ks,T', [(5000, 3, 50), (70000, 23, 300), (3000, 32, 40), (2000, 47, 90), (33, 16, 5),
                                      (100, 1, 7)])
def test_uh_vs_oracle(n, n_ks, T):
    rng = np.random.default_rng(n_ks)
    ker = rng.uniform(0, 1, (n_ks, n)) * (rng.random((n_ks, n)) < 0.55)
    lat = rng.gamma(0.3, 2e-3, (T, n)) * (rng.random((T, n)) < 0.4)
    s0 = rng.uniform(0, 1e-3, (n_ks, n))
    s_ref, s_gpu = s0.copy(), s0.copy()
    ref = oracle.uh_convolve(lat, ker, s_ref)
    out = uh_convolve(lat, ker, s_gpu)
    col = np.max(np.abs(ref), axis=0)
    assert parity_error(out, ref, col) < TOL
    assert parity_error(s_gpu, s_ref, col) < TOL
    assert np.all(s_gpu[-1] == 0)


def test_uh_inside_router_golden(route_golden):
    g = route_golden
    state = g['uh_state0'].copy()
    col = np.max(np.abs(g['conv']), axis=0) + np.max(np.abs(g['uh_state0']), axis=0)
    conv = uh_convolve(g['depths'], g['uh_kernel'], state)
    assert parity_error(conv, g['conv'], col) < TOL
    assert parity_error(state, g['uh_state1'], col) < TOL
    uh_convolve(g['depths2'], g['uh_kernel'], state)
    assert parity_error(state, g['uh_state2'], col) < TOL


@pytest.mark.parametrize('unit', ['m', 'mm'])
def test_weights_golden(unit):
    g = load_golden('weights.npz')
    indptr, indices, data = g[f'csr_indptr_{unit}'], g[f'csr_indices_{unit}'], g[f'csr_data_{unit}']
    for cumulative in (False, True):
        src = g['runoff_raw_cumulative'] if cumulative else g['runoff_raw']
        assert src.dtype == np.float32
        for vol in (False, True):
            ql = weights_transform(indptr, indices, data, src, cumulative=cumulative,
                                   area=g['catchment_area'] if vol else None)
            ref = g[f'ql_{unit}_cum{int(cumulative)}_vol{int(vol)}']
            assert not np.isnan(ql).any()
            assert parity_error(ql, ref) < TOL


def test_weights_irregular_time_axis_golden():
    """keep_nan + the host resample tail == the reference's runoff.py:316-337 on a grid with NaN cells."""
    from river_route_b200.runoff import resample_irregular
    g, gi = load_golden('weights.npz'), load_golden('weights_irregular.npz')
    indptr, indices, data = g['csr_indptr_m'], g['csr_indices_m'], g['csr_data_m']
    t_in = gi['time_index'].astype('datetime64[s]')
    for cumulative in (False, True):
        src = gi['runoff_raw_cumulative'] if cumulative else gi['runoff_raw']
        for vol in (False, True):
            ql = weights_transform(indptr, indices, data, src, cumulative=cumulative, keep_nan=True)
            assert np.isnan(ql).any()
            out, _ = resample_irregular(ql, t_in, g['river_ids_ordered'], g['catchment_area'] if vol else None)
            ref = gi[f'ql_cum{int(cumulative)}_vol{int(vol)}']
            assert parity_error(out, ref) < TOL
            assert np.array_equal(out == 0.0, ref == 0.0)


@pytest.mark.parametrize('f32', [True, False])
def test_weights_vs_oracle(f32):
    rng = np.random.default_rng(31)
    n_riv, n_pts, T = 60000, 40000, 37
    nnz_per = rng.integers(0, 13, n_riv)            # includes empty rows and rows longer than the register cache
    river_idx = np.repeat(np.arange(n_riv), nnz_per)
    point_idx = rng.integers(0, n_pts, river_idx.shape[0])
    vals = rng.random(river_idx.shape[0])
    indptr, indices, data = oracle.weights_csr(river_idx, point_idx, vals, n_riv, n_pts)
    x = rng.gamma(0.3, 2e-3, (T, n_pts))
    x[rng.random(x.shape) < 0.6] = 0.0
    x[5, 17] = np.nan
    x = np.cumsum(x, axis=0)
    if f32:
        x = x.astype(np.float32)
    area = rng.uniform(1e5, 5e8, n_riv)
    for cumulative in (False, True):
        for fp in (False, True):
            ref = oracle.weights_transform(indptr, indices, data, x, cumulative=cumulative, force_positive=fp, area=area)
            out = weights_transform(indptr, indices, data, x, cumulative=cumulative, force_positive=fp, area=area)
            assert parity_error(out, ref) < TOL
