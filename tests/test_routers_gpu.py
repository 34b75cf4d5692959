"""
The router classes end to end on the GPU, driven like the reference's own tests drive river_route:
real params / state parquet files and kernel npz, ``route()``, outputs captured through
``set_write_discharges`` (netCDF4 / xarray are not installed in this image, so lateral inflow is injected by
overriding ``_qlateral_generator`` -- the same seam SURVEY.md 8c uses to run the reference in memory).
Expected values are the reference's own outputs (tests/golden/).
"""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse

import river_route_b200 as rr
from tests.conftest import require_cuda
from tests.helpers import parity_error

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True)
def _cuda():
    require_cuda()


def _files(g, tmp_path):
    ids, down = g['river_ids'], g['down']
    params = str(tmp_path / 'params.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': g['k'], 'x': g['x']}).to_parquet(params)
    state = str(tmp_path / 'state.parquet')
    pd.DataFrame({'Q': g['q0']}).to_parquet(state)
    return params, state


def _dates(T, dt):
    return (np.datetime64('2020-01-01T00:00:00') + np.arange(T) * np.timedelta64(int(dt), 's')).astype('datetime64[s]')


class Capture:
    def __init__(self):
        self.calls = []

    def __call__(self, dates, q_array, q_file, routed_file=''):
        self.calls.append((dates.copy(), q_array.copy(), q_file, routed_file))


def _inject(cls, series, dt_runoff):
    class Injected(cls):
        def _qlateral_generator(self):
            for arr, out in zip(series, self.cfg.discharge_files):
                yield _dates(arr.shape[0], dt_runoff), arr, 'memory', out
    return Injected


def test_rapid_muskingum_route(route_golden, tmp_path):
    g = route_golden
    params, state = _files(g, tmp_path)
    cap = Capture()
    final = str(tmp_path / 'final.parquet')
    r = _inject(rr.RapidMuskingum, [g['ql']], g['dt_runoff'])(
        params_file=params, qlateral_files=[params], discharge_files=[str(tmp_path / 'q.nc')],
        channel_state_init_file=state, channel_state_final_file=final, dt_routing=int(g['dt_routing']), log=False)
    r.set_write_discharges(cap).route()
    dates, q, q_file, _ = cap.calls[0]
    assert q.dtype == np.float32 and q.shape == g['rapid_out'].shape          # TransformMuskingum.py:141-143
    np.testing.assert_allclose(q, g['rapid_out'].astype(np.float32), rtol=1e-6, atol=1e-6 * g['rapid_out'].max())
    assert parity_error(r.channel_state, g['rapid_q']) < TOL
    assert parity_error(pd.read_parquet(final)['Q'].values, g['rapid_q']) < TOL
    assert np.array_equal(r.river_ids, g['river_ids']) and np.array_equal(r.c1, g['c1']) and np.array_equal(r.c3, g['c3'])
    assert np.array_equal(r.A.toarray()[g['down'][g['down'] >= 0], np.flatnonzero(g['down'] >= 0)], np.ones((g['down'] >= 0).sum()))


def test_muskingum_route(route_golden, tmp_path):
    g = route_golden
    params, state = _files(g, tmp_path)
    cap = Capture()
    dt = int(g['dt_routing'])
    nrpo, nout = int(g['musk_nrpo']), int(g['musk_nout'])
    r = rr.Muskingum(params_file=params, discharge_dir=str(tmp_path), channel_state_init_file=state, dt_routing=dt,
                     dt_discharge=dt * nrpo, dt_total=dt * nrpo * nout, start_datetime='2021-03-01', log=False)
    r.set_write_discharges(cap).route()
    dates, q, q_file, _ = cap.calls[0]
    assert q_file.endswith('discharge.nc') and dates[0] == np.datetime64('2021-03-01')
    assert dates.shape[0] == nout and (dates[1] - dates[0]) == np.timedelta64(dt * nrpo, 's')
    np.testing.assert_allclose(q, g['musk_out'].astype(np.float32), rtol=1e-6, atol=1e-6 * g['musk_out'].max())
    assert parity_error(r.channel_state, g['musk_q']) < TOL
    assert (q >= 0).all()


def test_unit_muskingum_two_files_sequential(route_golden, tmp_path):
    """UH carry-over and channel state chain across files (reference: tests/test_unit_muskingum.py:75-146)."""
    g = route_golden
    params, state = _files(g, tmp_path)
    kfile = str(tmp_path / 'uh.npz')
    scipy.sparse.save_npz(kfile, scipy.sparse.csr_matrix(g['uh_kernel']))
    uh0 = str(tmp_path / 'uh0.parquet')
    pd.DataFrame(g['uh_state0'].T).to_parquet(uh0)                 # (n_basins, n_kernel_steps), UnitHydrograph.py:47-62
    uh1 = str(tmp_path / 'uh1.parquet')
    cap = Capture()
    r = _inject(rr.UnitMuskingum, [g['depths'], g['depths2']], g['dt_runoff'])(
        params_file=params, qlateral_files=[params, params],
        discharge_files=[str(tmp_path / 'a.nc'), str(tmp_path / 'b.nc')], channel_state_init_file=state,
        uh_kernel_file=kfile, uh_state_init_file=uh0, uh_state_final_file=uh1, dt_routing=int(g['dt_routing']), log=False)
    r.set_write_discharges(cap).route()
    assert len(cap.calls) == 2
    for (_, q, _, _), ref in zip(cap.calls, (g['unit_out'], g['unit_out2'])):
        np.testing.assert_allclose(q, ref.astype(np.float32), rtol=2e-6, atol=2e-6 * ref.max())
    assert parity_error(r.channel_state, g['unit_q2']) < TOL
    col = np.max(np.abs(g['conv']), axis=0) + np.max(np.abs(g['uh_state0']), axis=0)
    assert parity_error(pd.read_parquet(uh1).T.to_numpy(), g['uh_state2'], col) < TOL
    assert np.array_equal(r.hw_idx, g['hw_idx']) and np.array_equal(r.inner_idx, g['inner_idx'])


def test_split_run_equals_single_run_and_ensemble_mean(route_golden, tmp_path):
    """Two files at once == two runs chained through the state file (tests/test_rapid_muskingum.py:95-143);
    ensemble mode starts every member from the same state and ends with the member mean (:121-126, :145-146)."""
    g = route_golden
    params, state = _files(g, tmp_path)
    T = g['ql'].shape[0]
    a, b = g['ql'][: T // 2], g['ql'][T // 2: 2 * (T // 2)]
    common = dict(params_file=params, channel_state_init_file=state, dt_routing=int(g['dt_routing']), log=False)
    both = Capture()
    _inject(rr.RapidMuskingum, [a, b], g['dt_runoff'])(
        **common, qlateral_files=[params, params], discharge_files=[str(tmp_path / 'a.nc'), str(tmp_path / 'b.nc')]
    ).set_write_discharges(both).route()
    mid = str(tmp_path / 'mid.parquet')
    first, second = Capture(), Capture()
    _inject(rr.RapidMuskingum, [a], g['dt_runoff'])(
        **common, qlateral_files=[params], discharge_files=[str(tmp_path / 'a.nc')], channel_state_final_file=mid
    ).set_write_discharges(first).route()
    _inject(rr.RapidMuskingum, [b], g['dt_runoff'])(
        **dict(common, channel_state_init_file=mid), qlateral_files=[params], discharge_files=[str(tmp_path / 'b.nc')]
    ).set_write_discharges(second).route()
    assert np.array_equal(both.calls[0][1], first.calls[0][1]) and np.array_equal(both.calls[1][1], second.calls[0][1])
    ens = Capture()
    r = _inject(rr.RapidMuskingum, [a, b], g['dt_runoff'])(
        **common, qlateral_files=[params, params], discharge_files=[str(tmp_path / 'a.nc'), str(tmp_path / 'b.nc')],
        runoff_processing_mode='ensemble').set_write_discharges(ens)
    r.route()
    assert np.array_equal(ens.calls[0][1], first.calls[0][1])         # member 0 == sequential file 0
    assert not np.array_equal(ens.calls[1][1], second.calls[0][1])    # member 1 restarts from the initial state
    assert np.array_equal(r.channel_state, np.array(r._ensemble_member_states).mean(axis=0))


def test_resample_to_dt_discharge(route_golden, tmp_path):
    g = route_golden
    params, state = _files(g, tmp_path)
    T = (g['ql'].shape[0] // 3) * 3
    cap = Capture()
    _inject(rr.RapidMuskingum, [g['ql'][:T]], g['dt_runoff'])(
        params_file=params, qlateral_files=[params], discharge_files=[str(tmp_path / 'q.nc')],
        channel_state_init_file=state, dt_routing=int(g['dt_routing']), dt_discharge=3 * int(g['dt_runoff']), log=False
    ).set_write_discharges(cap).route()
    dates, q, _, _ = cap.calls[0]
    ref = g['rapid_out'][:T].reshape(T // 3, 3, -1).mean(axis=1)      # TransformMuskingum.py:128-139
    assert q.shape == ref.shape and dates.shape[0] == T // 3
    np.testing.assert_allclose(q, ref.astype(np.float32), rtol=1e-6, atol=1e-6 * ref.max())


def test_weights_to_qlateral_and_unit_hydrograph_classes():
    from tests.conftest import load_golden
    g = load_golden('weights.npz')
    table = {k: g[k] for k in ('river_id', 'x_index', 'y_index', 'proportion', 'area_sqm')}
    for unit in ('m', 'mm'):
        ql, rivers = rr.weights_to_qlateral(table, g['grid'], runoff_depth_unit=unit, as_volumes=True)
        assert np.array_equal(rivers, g['river_ids_ordered'])
        assert parity_error(ql, g[f'ql_{unit}_cum0_vol1']) < TOL
    with pytest.raises(ValueError, match='Unknown units'):
        rr.weights_to_qlateral(table, g['grid'], runoff_depth_unit='inches')
    u = load_golden('uh.npz')
    full, inc = rr.UnitHydrograph(kernel=u['kernel']), rr.UnitHydrograph(kernel=u['kernel'])
    res_full = full.convolve(u['lateral'])
    res_inc = np.array([inc.convolve_incrementally(row) for row in u['lateral']])
    np.testing.assert_allclose(res_full, res_inc, rtol=1e-12)         # tests/test_uhkernels.py:52-78
    np.testing.assert_allclose(res_full, u['conv_full'], rtol=1e-12)
    np.testing.assert_allclose(full.state, inc.state, rtol=1e-12, atol=1e-300)
    with pytest.raises(ValueError, match='does not match kernel shape'):
        import tempfile, os
        with tempfile.TemporaryDirectory() as d:
            pd.DataFrame(np.zeros((2, 2))).to_parquet(os.path.join(d, 's.parquet'))
            full.set_state(os.path.join(d, 's.parquet'))
