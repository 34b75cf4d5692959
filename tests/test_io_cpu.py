"""
Host-side file handling of the routers (no GPU): netCDF layouts of docs/references/io-file-schema.md read and written
through river_route_b200.ncio (netCDF4 when installed, classic netCDF-3 via scipy otherwise), the weight-table and
grid gathers that feed the device, and the rule that decides when the fp64 host intermediate may be skipped.
"""
import numpy as np
import pandas as pd
import pytest

import river_route_b200 as rr
from river_route_b200 import ncio, routers
from river_route_b200.runoff import (QlateralDataset, build_weight_csr, gather_grid_runoff, grid_runoff_unit,
                                     read_weight_table)
from river_route_b200.transforms import Transform
from oracle import oracle


def write_grid(path, ro, t0='2020-01-01 00:00:00', dt_hours=1, units='m', dims=('time', 'lat', 'lon'), time_units=None):
    """ERA5-like runoff file: ro(time, lat, lon) float32 with a CF time axis."""
    T, ny, nx = ro.shape
    with ncio.open_nc(path, 'w') as ds:
        ds.createDimension('time', T)
        ds.createDimension('lat', ny)
        ds.createDimension('lon', nx)
        tv = ds.createVariable('time', 'f8', ('time',))
        tv.units = time_units or f'hours since {t0}'
        tv[:] = np.arange(T, dtype=np.float64) * dt_hours
        la = ds.createVariable('lat', 'f8', ('lat',))
        la[:] = np.linspace(90, -90, ny)
        lo = ds.createVariable('lon', 'f8', ('lon',))
        lo[:] = np.linspace(0, 359, nx)
        order = [('time', 'lat', 'lon').index(d) for d in dims]
        v = ds.createVariable('ro', ro.dtype.str[1:], dims)
        if units is not None:
            v.units = units
        v[:] = np.transpose(ro, order)


def write_weight_table(path, table, var_river_id='river_id'):
    n = len(table['river_id'])
    with ncio.open_nc(path, 'w') as ds:
        ds.createDimension('index', n)
        for name, key, typ in ((var_river_id, 'river_id', 'i4'), ('x_index', 'x_index', 'i4'), ('y_index', 'y_index', 'i4'),
                               ('proportion', 'proportion', 'f8'), ('area_sqm', 'area_sqm', 'f8')):
            v = ds.createVariable(name, typ, ('index',))
            v[:] = np.asarray(table[key]).astype(typ)


def synthetic_table(n_rivers, ny, nx, seed=0, ids=None):
    rng = np.random.default_rng(seed)
    per = rng.integers(1, 5, n_rivers)
    ids = np.arange(1, n_rivers + 1) if ids is None else ids
    river = np.repeat(ids, per)
    m = river.shape[0]
    table = dict(river_id=river, x_index=rng.integers(0, nx, m), y_index=rng.integers(0, ny, m),
                 proportion=rng.uniform(0.05, 1.0, m), area_sqm=rng.uniform(1e5, 5e8, m))
    table['x_index'][1], table['y_index'][1] = table['x_index'][0], table['y_index'][0]   # a repeated cell
    return table


def test_decode_time_units():
    t = ncio.decode_time([0, 1.5, 3], 'hours since 1900-01-01 00:00:00.0')
    assert t.dtype == np.dtype('datetime64[s]') and t[1] - t[0] == np.timedelta64(5400, 's')
    assert t[0] == np.datetime64('1900-01-01T00:00:00')
    assert ncio.decode_time([86400], 'seconds since 2020-02-28')[0] == np.datetime64('2020-02-29T00:00:00')
    assert ncio.decode_time([2], 'days since 2001-01-01T12:00:00Z')[0] == np.datetime64('2001-01-03T12:00:00')
    with pytest.raises(ValueError, match='Unsupported time units'):
        ncio.decode_time([0], 'fortnights since 2000-01-01')


def test_discharge_file_layout_round_trip(tmp_path):
    """Muskingum.py:337-351: time f8 'seconds since', river id i4, Q f4 (time, river_id) + attributes."""
    ids = np.array([11, 12, 17], dtype=np.int64)
    pd.DataFrame({'river_id': ids, 'downstream_river_id': [12, 17, -1], 'k': [3000., 4000., 5000.],
                  'x': [0.2, 0.2, 0.3]}).to_parquet(tmp_path / 'p.parquet')
    r = rr.Muskingum(params_file=str(tmp_path / 'p.parquet'), discharge_dir=str(tmp_path), dt_routing=600,
                     dt_total=3600, channel_state_init_file=str(tmp_path / 'p.parquet'), log=False)
    r.river_ids = ids
    dates = (np.datetime64('2021-03-01T06:00:00') + np.arange(4) * np.timedelta64(3, 'h')).astype('datetime64[s]')
    q = np.arange(12, dtype=np.float32).reshape(4, 3) / 7
    r._write_discharges(dates, q, str(tmp_path / 'q.nc'), routed_file='runoff.nc')
    with ncio.open_nc(tmp_path / 'q.nc') as ds:
        assert tuple(ds.variables['Q'].dimensions) == ('time', 'river_id')
        Q = ncio.read_array(ds.variables['Q'])
        assert Q.dtype == np.float32 and np.array_equal(Q, q)
        rid = ncio.read_array(ds.variables['river_id'])
        assert rid.dtype == np.int32 and np.array_equal(rid, ids)
        tv = ds.variables['time']
        assert ncio.attrs_of(tv)['units'] == 'seconds since 2021-03-01 06:00:00'
        assert np.array_equal(ncio.decode_time(ncio.read_array(tv), ncio.attrs_of(tv)['units']), dates)
        qa = ncio.attrs_of(ds.variables['Q'])
        assert qa['units'] == 'm3 s-1' and qa['aggregation_method'] == 'mean' and qa['standard_name'] == 'discharge'
        assert (ds.runoff_file.decode() if isinstance(ds.runoff_file, bytes) else ds.runoff_file) == 'runoff.nc'


def test_qlateral_file_round_trip(tmp_path):
    ql = np.random.default_rng(0).uniform(0, 5, (6, 4))
    t = (np.datetime64('2020-01-01') + np.arange(6) * np.timedelta64(3, 'h')).astype('datetime64[s]')
    QlateralDataset(ql, [5, 6, 7, 9], t, 'm3').to_netcdf(str(tmp_path / 'ql.nc'))
    dates, arr = routers._read_qlateral(str(tmp_path / 'ql.nc'))
    assert np.array_equal(dates, t) and arr.dtype == np.float64 and np.array_equal(arr, ql)


@pytest.mark.parametrize('dims', [('time', 'lat', 'lon'), ('time', 'lon', 'lat'), ('lat', 'lon', 'time')])
def test_grid_gather_matches_pointwise_isel(tmp_path, dims):
    """runoff.py:267-280: ro.isel(x=cells_x, y=cells_y) -> (time, points), whatever the file's dimension order."""
    rng = np.random.default_rng(1)
    ro = rng.gamma(0.3, 2e-3, (70, 9, 13)).astype(np.float32)
    write_grid(str(tmp_path / 'g.nc'), ro, units='mm', dims=dims)
    table = synthetic_table(40, 9, 13)
    write_weight_table(str(tmp_path / 'w.nc'), table)
    back = read_weight_table(str(tmp_path / 'w.nc'))
    for key in table:
        assert np.array_equal(back[key], table[key]) if key.endswith('index') or key == 'river_id' \
            else np.allclose(back[key], table[key], rtol=0, atol=0)
    _, _, _, cx, cy, rivers, _ = build_weight_csr(back['river_id'], back['x_index'], back['y_index'], back['proportion'],
                                                  back['area_sqm'])
    dates, raw = gather_grid_runoff(str(tmp_path / 'g.nc'), cx, cy, var_x='lon', var_y='lat', slab_rows=16)
    assert raw.dtype == np.float32 and raw.flags.c_contiguous and np.array_equal(raw, ro[:, cy, cx])
    assert dates[0] == np.datetime64('2020-01-01T00:00:00') and dates[1] - dates[0] == np.timedelta64(3600, 's')
    assert grid_runoff_unit(str(tmp_path / 'g.nc')) == 'mm'
    # two consecutive files concatenate along time like xr.open_mfdataset
    write_grid(str(tmp_path / 'g2.nc'), ro[:5], t0='2020-01-03 22:00:00', dims=dims)
    d2, raw2 = gather_grid_runoff([str(tmp_path / 'g.nc'), str(tmp_path / 'g2.nc')], cx, cy, var_x='lon', var_y='lat')
    assert raw2.shape[0] == 75 and np.array_equal(raw2[70:], ro[:5][:, cy, cx]) and d2.shape[0] == 75
    write_grid(str(tmp_path / 'g3.nc'), ro[:2], units=None)
    assert grid_runoff_unit(str(tmp_path / 'g3.nc')) == 'm'                       # runoff.py:268 default
    # flat layout for the device-side gather: the whole (time, y, x) grid with columns = flat cell ids y * nx + x
    from river_route_b200.runoff import grid_layout
    assert grid_layout(str(tmp_path / 'g.nc'), var_x='lon', var_y='lat') == (dims, 9, 13)
    if dims == ('time', 'lat', 'lon'):
        _, flat = gather_grid_runoff(str(tmp_path / 'g.nc'), cx, cy, var_x='lon', var_y='lat', flat=True, slab_rows=32)
        assert flat.shape == (70, 9 * 13) and np.array_equal(flat[:, cy * 13 + cx], raw)
    else:
        with pytest.raises(ValueError, match='stored as .time, y, x.'):
            gather_grid_runoff(str(tmp_path / 'g.nc'), cx, cy, var_x='lon', var_y='lat', flat=True)


def test_weight_csr_equals_scipy_construction():
    """build_weight_csr == scipy's COO -> CSR of runoff.py:283-295 (duplicates summed, columns ascending) and the
    oracle's restatement; rivers and cells are numbered by first appearance."""
    import scipy.sparse
    table = synthetic_table(300, 20, 30, seed=3, ids=np.random.default_rng(3).permutation(300) + 1000)
    indptr, indices, data, cx, cy, rivers, area = build_weight_csr(
        table['river_id'], table['x_index'], table['y_index'], table['proportion'], table['area_sqm'], 0.001)
    df = pd.DataFrame(table)
    uniq = df[['x_index', 'y_index']].drop_duplicates().reset_index(drop=True).reset_index()
    pidx = df[['x_index', 'y_index']].merge(uniq, on=['x_index', 'y_index'], how='left')['index'].values
    rid = df[['river_id']].drop_duplicates()['river_id'].values
    ridx = pd.Series(np.arange(len(rid)), index=rid).loc[df['river_id'].values].values
    W = scipy.sparse.csr_matrix((df['proportion'].values * 0.001, (ridx, pidx)), shape=(len(rid), len(uniq)))
    W.sum_duplicates()
    W.sort_indices()
    assert np.array_equal(indptr, W.indptr) and np.array_equal(indices, W.indices) and np.array_equal(data, W.data)
    assert np.array_equal(rivers, rid) and np.array_equal(cx, uniq['x_index'].values) and np.array_equal(cy, uniq['y_index'].values)
    o_ptr, o_idx, o_dat = oracle.weights_csr(ridx, pidx, df['proportion'].values * 0.001, len(rid), len(uniq))
    assert np.array_equal(o_ptr, indptr) and np.array_equal(o_idx, indices) and np.array_equal(o_dat, data)
    assert np.allclose(area, df.groupby(ridx)['area_sqm'].sum().values, rtol=0, atol=0)


def test_stock_router_detection():
    """The fp64 host intermediate is skipped only when nobody can observe it: ``_router`` / ``_qlateral_generator``
    not overridden by a subclass nor patched on the instance (SURVEY.md 8b seams)."""
    class Sub(rr.RapidMuskingum):
        def _router(self, qlateral):
            return super()._router(qlateral)

    class Gen(rr.UnitMuskingum):
        def _qlateral_generator(self):
            yield from ()

    def make(cls):
        return cls.__new__(cls)

    T = routers.TransformMuskingum
    assert routers._is_stock(make(rr.RapidMuskingum), '_router', T)
    assert routers._is_stock(make(rr.UnitMuskingum), '_router', T)
    assert not routers._is_stock(make(Sub), '_router', T)
    assert routers._is_stock(make(Gen), '_router', T) and not routers._is_stock(make(Gen), '_qlateral_generator', T)
    patched = make(rr.RapidMuskingum)
    patched._router = lambda ql: None
    assert not routers._is_stock(patched, '_router', T)
    assert routers._is_stock(make(rr.Muskingum), '_router', rr.Muskingum)


def test_transform_and_stream_arguments_fail_before_the_device():
    """Malformed weight tables and shapes raise on the host; a well-formed call without a GPU reports that there is
    no CPU fallback (status 201) instead of computing anything."""
    ok = (np.array([0, 1, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32), np.ones(2))
    with pytest.raises(RuntimeError, match='outside the gathered runoff array'):
        Transform(ok[0], np.array([0, 5], dtype=np.int32), ok[2], 3)
    with pytest.raises(RuntimeError, match='non-decreasing'):
        Transform(np.array([0, 2, 1], dtype=np.int32), ok[1], ok[2], 3)
    with pytest.raises(ValueError, match='one value per river segment'):
        Transform(*ok, 3, area=np.ones(5))
    plan = rr.Plan(np.array([1, -1], dtype=np.int32))
    with pytest.raises(ValueError, match='does not match'):
        plan.route_host(rr.MODE_RAPID, np.zeros(2), np.zeros((5, 2)), np.empty((2, 2), dtype=np.float32), 1, resample=2)
    with pytest.raises(ValueError, match='float64 .or float32.'):
        plan.route_host(rr.MODE_RAPID, np.zeros(2), np.zeros((4, 2)), np.empty((4, 2), dtype=np.int32), 1)
    if not rr.cuda_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            Transform(*ok, 3)
        plan.set_coefficients(np.full(2, .2), np.full(2, .3), np.full(2, .5), np.full(2, 1e-4))
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            plan.route_host(rr.MODE_RAPID, np.zeros(2), np.zeros((4, 2)), np.empty((2, 2), dtype=np.float32), 1, resample=2)
    plan.close()
