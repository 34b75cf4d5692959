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


# ------------------------------------------------------------------------------------------------------
# file level: YAML config -> params parquet + grid runoff netCDF + weight table netCDF -> discharge netCDF
# ------------------------------------------------------------------------------------------------------
def _grid_case(tmp_path, n=3000, T=30, ny=15, nx=22, cumulative=False, units='m', dims=('time', 'lat', 'lon')):
    from river_route_b200 import synth
    from tests.test_io_cpu import synthetic_table, write_grid, write_weight_table
    down = synth.forest(n, 3, seed=21, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 21)
    ids = np.arange(n, dtype=np.int64) * 3 + 100
    params = str(tmp_path / 'params.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(params)
    rng = np.random.default_rng(5)
    q0 = rng.uniform(0, 20, n)
    state = str(tmp_path / 'state.parquet')
    pd.DataFrame({'Q': q0}).to_parquet(state)
    table = synthetic_table(n, ny, nx, seed=7, ids=ids)
    write_weight_table(str(tmp_path / 'weights.nc'), table)
    grids = []
    for f in range(2):
        ro = (rng.gamma(0.3, 2e-3, (T, ny, nx)) * (rng.random((T, ny, nx)) < 0.5)).astype(np.float32)
        if units == 'mm':
            ro *= 1000
        if cumulative:
            ro = np.cumsum(ro, axis=0, dtype=np.float32)
        path = str(tmp_path / f'runoff_{f}.nc')
        write_grid(path, ro, t0=f'2020-01-0{1 + f} 00:00:00', dt_hours=1, units=units, dims=dims)
        grids.append((path, ro))
    return dict(down=down, k=k, x=x, ids=ids, q0=q0, table=table, grids=grids, params=params, state=state, n=n, T=T)


def _oracle_grid_chain(c, dt, cumulative, units, unit_hydrograph=None):
    """The reference's sequence on the CPU oracle: runoff_to_qlateral -> [UH] -> route, file after file."""
    from oracle import oracle
    from river_route_b200.runoff import build_weight_csr
    from tests.helpers import network_arrays
    a = network_arrays(c['down'], c['k'], c['x'], dt, dt)
    t = c['table']
    indptr, indices, data, cx, cy, rivers, area = build_weight_csr(t['river_id'], t['x_index'], t['y_index'],
                                                                   t['proportion'], t['area_sqm'], 0.001 if units == 'mm' else 1)
    q, outs = c['q0'].copy(), []
    for _, ro in c['grids']:
        raw = ro[:, cy, cx]
        if unit_hydrograph is None:
            ql = oracle.weights_transform(indptr, indices, data, raw, cumulative=cumulative, area=area)
            out = np.zeros((c['T'], c['n']))
            oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, out, 1)
        else:
            ker, st = unit_hydrograph
            conv = oracle.uh_convolve(oracle.weights_transform(indptr, indices, data, raw, cumulative=cumulative), ker, st)
            sp = oracle.unit_split(c['down'].astype(np.int64))
            inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
            c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
            q_ch = q[inner].copy()
            q_full = q_ch.copy()
            out = np.zeros((c['T'], c['n']))
            oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw,
                              inner, q_ch, q_full, conv, out, 1)
            q = np.empty(c['n'])
            q[hw] = conv[-1][hw]
            q[inner] = q_full
        outs.append(out)
    return outs, q


@pytest.mark.parametrize('cumulative,units,dims', [(False, 'm', ('time', 'lat', 'lon')), (True, 'mm', ('time', 'lat', 'lon')),
                                                   (False, 'mm', ('lat', 'time', 'lon'))])
def test_rapid_muskingum_from_grid_files_yaml_config(tmp_path, cumulative, units, dims):
    """The fused device path behind the unchanged config surface (examples/config.yaml keys): nothing is injected.
    (time, y, x) files are shipped whole and gathered by the SpMM on the device (flat cell ids); other dimension
    orders are gathered on the host first."""
    import yaml
    from river_route_b200 import ncio
    c = _grid_case(tmp_path, cumulative=cumulative, units=units, dims=dims)
    out_dir = tmp_path / 'out'
    out_dir.mkdir()
    cfg = dict(params_file=c['params'], grid_runoff_files=[g[0] for g in c['grids']], grid_weights_file=str(tmp_path / 'weights.nc'),
               discharge_dir=str(out_dir), channel_state_init_file=c['state'], channel_state_final_file=str(tmp_path / 'final.parquet'),
               grid_accumulation_type='cumulative' if cumulative else 'incremental', var_x='lon', var_y='lat', log=False)
    with open(tmp_path / 'config.yaml', 'w') as f:
        yaml.safe_dump(cfg, f)
    r = rr.RapidMuskingum(str(tmp_path / 'config.yaml'))
    before = rr.launch_count()
    r.route()
    assert rr.launch_count() > before and r._transform is not None      # the device-resident path ran
    assert (r._transform_key[1] is not None) == (dims == ('time', 'lat', 'lon'))   # device-side gather only for (t, y, x)
    outs, q_final = _oracle_grid_chain(c, 3600, cumulative, units)
    for f, ref in enumerate(outs):
        with ncio.open_nc(out_dir / f'discharge_runoff_{f}.nc') as ds:
            Q = ncio.read_array(ds.variables['Q'])
            assert np.array_equal(ncio.read_array(ds.variables['river_id']), c['ids'].astype(np.int32))
            tv = ds.variables['time']
            dates = ncio.decode_time(ncio.read_array(tv), ncio.attrs_of(tv)['units'])
            assert dates[0] == np.datetime64(f'2020-01-0{1 + f}T00:00:00') and dates.shape[0] == c['T']
        assert Q.dtype == np.float32 and Q.shape == ref.shape
        np.testing.assert_allclose(Q, ref.astype(np.float32), rtol=2e-6, atol=2e-6 * ref.max())
    assert parity_error(pd.read_parquet(tmp_path / 'final.parquet')['Q'].values, q_final) < TOL
    # the unfused sequence (override the seam -> fp64 host arrays, as the reference does it) writes the same bytes
    class Seam(rr.RapidMuskingum):
        def _router(self, qlateral):
            return super()._router(qlateral)
    cap = Capture()
    Seam(**dict(cfg, channel_state_final_file=None)).set_write_discharges(cap).route()
    for f in range(2):
        with ncio.open_nc(out_dir / f'discharge_runoff_{f}.nc') as ds:
            assert np.array_equal(ncio.read_array(ds.variables['Q']), cap.calls[f][1])


def test_unit_muskingum_from_grid_files(tmp_path):
    c = _grid_case(tmp_path, n=2500, T=20)
    rng = np.random.default_rng(9)
    ker = rng.uniform(0, 1, (7, c['n'])) * (rng.random((7, c['n'])) < 0.7)
    kfile = str(tmp_path / 'uh.npz')
    scipy.sparse.save_npz(kfile, scipy.sparse.csr_matrix(ker))
    cap = Capture()
    uh1 = str(tmp_path / 'uh1.parquet')
    r = rr.UnitMuskingum(params_file=c['params'], grid_runoff_files=[g[0] for g in c['grids']],
                         grid_weights_file=str(tmp_path / 'weights.nc'), discharge_dir=str(tmp_path),
                         channel_state_init_file=c['state'], uh_kernel_file=kfile, uh_state_final_file=uh1,
                         var_x='lon', var_y='lat', dt_discharge=7200, log=False)
    r.set_write_discharges(cap).route()
    st = np.zeros_like(ker)
    outs, q_final = _oracle_grid_chain(c, 3600, False, 'm', unit_hydrograph=(ker, st))
    for (dates, q, _, _), ref in zip(cap.calls, outs):
        ref2 = ref.reshape(c['T'] // 2, 2, -1).mean(axis=1)
        assert q.shape == ref2.shape and dates.shape[0] == c['T'] // 2 and dates[1] - dates[0] == np.timedelta64(7200, 's')
        np.testing.assert_allclose(q, ref2.astype(np.float32), rtol=2e-6, atol=2e-6 * ref2.max())
    assert parity_error(r.channel_state, q_final) < TOL
    col = np.max(np.abs(st), axis=0) + 1e-300
    assert parity_error(pd.read_parquet(uh1).T.to_numpy(), st, col) < TOL


def test_qlateral_files_on_disk(route_golden, tmp_path):
    """qlateral netCDF in, discharge netCDF out -- the reference's basic RapidMuskingum run (test_rapid_muskingum.py)."""
    from river_route_b200 import ncio
    from river_route_b200.runoff import QlateralDataset
    g = route_golden
    params, state = _files(g, tmp_path)
    QlateralDataset(g['ql'], g['river_ids'], _dates(g['ql'].shape[0], g['dt_runoff']), 'm3').to_netcdf(str(tmp_path / 'ql.nc'))
    rr.RapidMuskingum(params_file=params, qlateral_files=str(tmp_path / 'ql.nc'), discharge_dir=str(tmp_path),
                      channel_state_init_file=state, dt_routing=int(g['dt_routing']), log=False).route()
    with ncio.open_nc(tmp_path / 'discharge_ql.nc') as ds:
        Q = ncio.read_array(ds.variables['Q'])
    np.testing.assert_allclose(Q, g['rapid_out'].astype(np.float32), rtol=1e-6, atol=1e-6 * g['rapid_out'].max())


def test_initial_state_effect_decays(route_golden, tmp_path):
    """tests/test_rapid_muskingum.py:46-92 of the reference: a different initial state changes the first step and
    its influence decays along the run."""
    g = route_golden
    params, state = _files(g, tmp_path)
    hot = str(tmp_path / 'hot.parquet')
    pd.DataFrame({'Q': g['q0'] + 100.0}).to_parquet(hot)
    T = g['ql'].shape[0]
    long_ql = np.concatenate([g['ql']] * 6)                                   # long enough for the state to flush
    runs = []
    for st in (state, hot):
        cap = Capture()
        _inject(rr.RapidMuskingum, [long_ql], g['dt_runoff'])(
            params_file=params, qlateral_files=[params], discharge_files=[str(tmp_path / 'q.nc')],
            channel_state_init_file=st, dt_routing=int(g['dt_routing']), log=False).set_write_discharges(cap).route()
        runs.append(cap.calls[0][1].astype(np.float64))
    diff = np.abs(runs[1] - runs[0]).mean(axis=1)
    assert diff[0] > 1.0                                                      # the first step sees the extra water
    assert diff[-1] < 0.6 * diff[0] and diff[-1] < diff[T]                    # ... and it drains away


def test_cumulative_runoff_equals_incremental():
    """tests/test_runoff.py:67-99 of the reference: cumulative input differenced on the device gives the same lateral
    inflows as the incremental input (up to the rounding of the running sum)."""
    from tests.conftest import load_golden
    g = load_golden('weights.npz')
    table = {k: g[k] for k in ('river_id', 'x_index', 'y_index', 'proportion', 'area_sqm')}
    inc = np.abs(g['grid']).astype(np.float64)
    cum = np.cumsum(inc, axis=0)
    a, _ = rr.weights_to_qlateral(table, inc, as_volumes=True)
    b, _ = rr.weights_to_qlateral(table, cum, cumulative=True, as_volumes=True)
    np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-9 * np.abs(a).max())
    c, _ = rr.weights_to_qlateral(table, -inc, force_positive_runoff=True)
    assert np.all(c == 0)                                                     # runoff.py:313-314


def test_output_subset_of_rivers(route_golden, tmp_path):
    """``set_output_rivers``: the device-side form of the reference's subset writer (docs/tutorial/advanced.md:147-170):
    the writer sees exactly the columns of the full run, the state is the full state, unknown ids are rejected."""
    from river_route_b200 import ncio
    g = route_golden
    params, state = _files(g, tmp_path)
    common = dict(params_file=params, qlateral_files=[params], channel_state_init_file=state,
                  dt_routing=int(g['dt_routing']), log=False)
    full = Capture()
    r0 = _inject(rr.RapidMuskingum, [g['ql']], g['dt_runoff'])(**common, discharge_files=[str(tmp_path / 'full.nc')])
    r0.set_write_discharges(full).route()
    ids = g['river_ids']
    pick = ids[[len(ids) - 1, 0, len(ids) // 2, 0]]                     # any order, repeats allowed
    r1 = _inject(rr.RapidMuskingum, [g['ql']], g['dt_runoff'])(**common, discharge_files=[str(tmp_path / 'sub.nc')])
    r1.set_output_rivers(pick).route()                                  # default netCDF writer
    with ncio.open_nc(tmp_path / 'sub.nc') as ds:
        Q = ncio.read_array(ds.variables['Q'])
        assert np.array_equal(ncio.read_array(ds.variables['river_id']), pick.astype(np.int32))
    cols = [int(np.flatnonzero(ids == p)[0]) for p in pick]
    assert Q.shape == (g['ql'].shape[0], 4) and np.array_equal(Q, full.calls[0][1][:, cols])
    assert np.array_equal(r1.channel_state, r0.channel_state)

    class Seam(rr.RapidMuskingum):                                      # overridden seam -> subset taken on the host
        def _router(self, qlateral):
            return super()._router(qlateral)
    cap = Capture()
    _inject(Seam, [g['ql']], g['dt_runoff'])(**common, discharge_files=[str(tmp_path / 's2.nc')]) \
        .set_output_rivers(pick).set_write_discharges(cap).route()
    assert np.array_equal(cap.calls[0][1], Q)
    with pytest.raises(ValueError, match='ids not in the params file'):
        _inject(rr.RapidMuskingum, [g['ql']], g['dt_runoff'])(**common, discharge_files=[str(tmp_path / 'x.nc')]) \
            .set_output_rivers([int(ids.max()) + 12345]).route()


def _write_ql(path, ql, ids, t0_hours=0):
    from river_route_b200 import ncio
    T, n = ql.shape
    with ncio.open_nc(path, 'w') as nc:
        nc.createDimension('time', T)
        nc.createDimension('river_id', n)
        tv = nc.createVariable('time', 'f8', ('time',))
        tv.units = 'seconds since 2022-01-01 00:00:00'
        tv[:] = (np.arange(T) + t0_hours) * 3600.0
        nc.createVariable('river_id', 'i4', ('river_id',))[:] = ids.astype(np.int32)
        nc.createVariable('qlateral', 'f4' if ql.dtype == np.float32 else 'f8', ('time', 'river_id'))[:] = ql


def _stream_case(tmp_path, n=5000, T=40, files=2, f32=False, seed=3):
    from river_route_b200 import synth
    down = synth.forest(n, 6, seed=seed, depth_bias=0.6)
    k, x = synth.muskingum_params(n, seed)
    ids = np.arange(n, dtype=np.int64) + 1000
    params = str(tmp_path / 'p.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(params)
    q0 = np.random.default_rng(seed).uniform(0, 30, n)
    pd.DataFrame({'Q': q0}).to_parquet(tmp_path / 'q0.parquet')
    paths, laterals = [], []
    for f in range(files):
        ql = synth.lateral_volumes(T, n, 80 + f)
        if f32:
            ql = ql.astype(np.float32)
        _write_ql(str(tmp_path / f'ql_{f}.nc'), ql, ids, f * T)
        paths.append(str(tmp_path / f'ql_{f}.nc'))
        laterals.append(ql)
    return dict(params=params, state=str(tmp_path / 'q0.parquet'), files=paths, laterals=laterals, down=down, k=k, x=x, q0=q0,
                ids=ids, n=n, T=T)


def _read_q(path):
    from river_route_b200 import ncio
    with ncio.open_nc(path) as ds:
        return ncio.read_array(ds.variables['Q'])


@pytest.mark.parametrize('router,f32,k', [('rapid', False, 1), ('rapid', True, 2), ('unit', False, 1)])
def test_qlateral_files_streamed_through_pinned_slabs(tmp_path, monkeypatch, router, f32, k):
    """Slab streaming on the real library: 16-row slabs give the bytes of one whole-file slab (float32 variables routed as
    stored, resample on slab boundaries, channel state / UH carry-over chained through slabs and files), and the
    float32 discharge agrees with the CPU oracle."""
    from oracle import oracle
    from tests.helpers import network_arrays
    c = _stream_case(tmp_path, f32=f32)
    scale = 1.0 if router == 'rapid' else 1e-7
    if router == 'unit':
        for f, ql in enumerate(c['laterals']):
            c['laterals'][f] = ql * scale
            _write_ql(c['files'][f], c['laterals'][f], c['ids'], f * c['T'])
        ker = np.random.default_rng(2).uniform(0, 1, (6, c['n'])) * (np.random.default_rng(3).random((6, c['n'])) < 0.7)
        scipy.sparse.save_npz(str(tmp_path / 'uh.npz'), scipy.sparse.csr_matrix(ker))
    outs = {}
    for label, rows in (('whole', None), ('slabs', '16')):
        out_dir = tmp_path / label
        out_dir.mkdir()
        if rows:
            monkeypatch.setenv('RR_ROUTER_SLAB_ROWS', rows)
        else:
            monkeypatch.delenv('RR_ROUTER_SLAB_ROWS', raising=False)
        cfg = dict(params_file=c['params'], qlateral_files=c['files'], discharge_dir=str(out_dir), dt_discharge=3600 * k,
                   channel_state_init_file=c['state'], log=False)
        r = (rr.UnitMuskingum(uh_kernel_file=str(tmp_path / 'uh.npz'), **cfg) if router == 'unit' else rr.RapidMuskingum(**cfg)).route()
        outs[label] = ([_read_q(out_dir / f'discharge_ql_{f}.nc') for f in range(2)], r.channel_state.copy())
    for f in range(2):
        assert np.array_equal(outs['whole'][0][f], outs['slabs'][0][f]), f
    assert np.array_equal(outs['whole'][1], outs['slabs'][1])
    if router == 'rapid':
        a = network_arrays(c['down'], c['k'], c['x'], 3600, 3600)
        q = c['q0'].copy()
        for f, ql in enumerate(c['laterals']):
            ref = np.zeros((c['T'], c['n']))
            oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql.astype(np.float64), ref, 1)
            want = (ref.reshape(c['T'] // k, k, c['n']).mean(axis=1) if k > 1 else ref).astype(np.float32)
            assert np.allclose(outs['slabs'][0][f], want, rtol=3e-7, atol=1e-30), f
        assert parity_error(outs['slabs'][1], q) < TOL


def test_ensemble_mode_batched_members_equal_single_member_runs(tmp_path, monkeypatch):
    """runoff_processing_mode='ensemble' on the real library: the batched device calls write, for every member, the bytes
    a single-member run writes, and the final state is numpy's member-order mean of the members' states."""
    c = _stream_case(tmp_path, n=4000, T=40, files=4, f32=True, seed=7)
    monkeypatch.setenv('RR_ROUTER_SLAB_ROWS', '16')
    ens_dir = tmp_path / 'ens'
    ens_dir.mkdir()
    r = rr.RapidMuskingum(params_file=c['params'], qlateral_files=c['files'], discharge_dir=str(ens_dir),
                          channel_state_init_file=c['state'], runoff_processing_mode='ensemble', log=False).route()
    states = []
    for f, path in enumerate(c['files']):
        one = tmp_path / f'one_{f}'
        one.mkdir()
        s = rr.RapidMuskingum(params_file=c['params'], qlateral_files=[path], discharge_dir=str(one),
                              channel_state_init_file=c['state'], log=False).route()
        states.append(s.channel_state.copy())
        assert np.array_equal(_read_q(ens_dir / f'discharge_ql_{f}.nc'), _read_q(one / f'discharge_ql_{f}.nc')), f
    assert np.array_equal(np.array(r._ensemble_member_states), np.array(states))
    assert np.array_equal(r.channel_state, np.array(states).mean(axis=0))
