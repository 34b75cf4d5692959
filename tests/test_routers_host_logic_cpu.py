"""
Host logic of the router classes without a GPU: the device calls (`Plan.route_host`, `Plan.runoff_route_host`,
`Transform`) are replaced by oracle-backed stand-ins with the same contracts, so that everything around them runs on
the CPU box -- YAML config -> params parquet / weight-table + grid / qlateral netCDF -> time bookkeeping -> fused or
unfused path selection -> device-side gather mapping (flat cell ids) -> output subset -> dt_discharge resample ->
state chaining across files -> UH carry-over sync -> discharge netCDF.  Expected values come from the plain oracle
chain on the gathered arrays (tests/test_routers_gpu.py::_oracle_grid_chain), i.e. from a different route through the
data than the one under test.  The GPU suite runs the same scenarios against the real library.
"""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse

import river_route_b200 as rr
from river_route_b200 import ncio, plan as plan_mod, routers, transforms
from oracle import oracle
from tests.helpers import network_arrays, parity_error
from tests import fakes
from tests.test_routers_gpu import Capture, _grid_case, _oracle_grid_chain


@pytest.fixture
def host_only(monkeypatch):
    """Oracle-backed stand-ins for the device entry points the routers use (tests/fakes.py)."""
    fakes.install(monkeypatch.setattr)


@pytest.mark.parametrize('cumulative,units,dims', [(False, 'm', ('time', 'lat', 'lon')), (True, 'mm', ('time', 'lat', 'lon')),
                                                   (False, 'mm', ('lon', 'time', 'lat'))])
def test_rapid_from_grid_files_yaml(tmp_path, host_only, cumulative, units, dims):
    import yaml
    c = _grid_case(tmp_path, n=600, T=12, ny=9, nx=14, cumulative=cumulative, units=units, dims=dims)
    out_dir = tmp_path / 'out'
    out_dir.mkdir()
    cfg = dict(params_file=c['params'], grid_runoff_files=[g[0] for g in c['grids']], grid_weights_file=str(tmp_path / 'weights.nc'),
               discharge_dir=str(out_dir), channel_state_init_file=c['state'], channel_state_final_file=str(tmp_path / 'final.parquet'),
               grid_accumulation_type='cumulative' if cumulative else 'incremental', var_x='lon', var_y='lat', log=False)
    with open(tmp_path / 'config.yaml', 'w') as f:
        yaml.safe_dump(cfg, f)
    r = rr.RapidMuskingum(str(tmp_path / 'config.yaml')).route()
    flat = dims == ('time', 'lat', 'lon')
    assert (r._transform_key[1] == (9, 14)) if flat else (r._transform_key[1] is None)
    assert r._transform.n_points == (9 * 14 if flat else len(r._cells[0]))
    outs, q_final = _oracle_grid_chain(c, 3600, cumulative, units)
    for f, ref in enumerate(outs):
        with ncio.open_nc(out_dir / f'discharge_runoff_{f}.nc') as ds:
            Q = ncio.read_array(ds.variables['Q'])
            tv = ds.variables['time']
            dates = ncio.decode_time(ncio.read_array(tv), ncio.attrs_of(tv)['units'])
        assert Q.dtype == np.float32 and np.array_equal(Q, ref.astype(np.float32))       # same arithmetic, other data path
        assert dates[0] == np.datetime64(f'2020-01-0{1 + f}T00:00:00') and dates.shape[0] == c['T']
    assert np.array_equal(pd.read_parquet(tmp_path / 'final.parquet')['Q'].values, q_final)
    # a weight table whose rivers are not the params file's rivers is refused instead of silently misrouted
    from tests.test_io_cpu import write_weight_table
    bad = dict(c['table'])
    bad['river_id'] = c['table']['river_id'][::-1].copy()
    write_weight_table(str(tmp_path / 'bad.nc'), bad)
    with pytest.raises(ValueError, match='same order'):
        rr.RapidMuskingum(**dict(cfg, grid_weights_file=str(tmp_path / 'bad.nc'))).route()


def test_unit_from_grid_files_resampled_with_state_files(tmp_path, host_only):
    c = _grid_case(tmp_path, n=500, T=10, ny=8, nx=11)
    rng = np.random.default_rng(9)
    ker = rng.uniform(0, 1, (5, c['n'])) * (rng.random((5, c['n'])) < 0.7)
    kfile = str(tmp_path / 'uh.npz')
    scipy.sparse.save_npz(kfile, scipy.sparse.csr_matrix(ker))
    s0 = rng.uniform(0, 1e-3, ker.shape)
    s0[-1] = 0
    pd.DataFrame(s0.T).to_parquet(tmp_path / 'uh0.parquet')
    cap = Capture()
    r = rr.UnitMuskingum(params_file=c['params'], grid_runoff_files=[g[0] for g in c['grids']],
                         grid_weights_file=str(tmp_path / 'weights.nc'), discharge_dir=str(tmp_path),
                         channel_state_init_file=c['state'], uh_kernel_file=kfile, uh_state_init_file=str(tmp_path / 'uh0.parquet'),
                         uh_state_final_file=str(tmp_path / 'uh1.parquet'), var_x='lon', var_y='lat', dt_discharge=7200, log=False)
    r.set_write_discharges(cap).route()
    st = s0.copy()
    outs, q_final = _oracle_grid_chain(c, 3600, False, 'm', unit_hydrograph=(ker, st))
    for (dates, q, _, _), ref in zip(cap.calls, outs):
        ref2 = ref.reshape(c['T'] // 2, 2, -1).mean(axis=1).astype(np.float32)
        assert np.array_equal(q, ref2) and dates.shape[0] == c['T'] // 2 and dates[1] - dates[0] == np.timedelta64(7200, 's')
    assert np.array_equal(r.channel_state, q_final)
    assert np.array_equal(pd.read_parquet(tmp_path / 'uh1.parquet').T.to_numpy(), st)   # carry-over came back from the "device"


def test_output_subset_and_overridden_seam(tmp_path, host_only):
    c = _grid_case(tmp_path, n=400, T=8, ny=7, nx=9)
    ids = c['ids']
    pick = ids[[399, 0, 200, 0]]
    cfg = dict(params_file=c['params'], grid_runoff_files=[c['grids'][0][0]], grid_weights_file=str(tmp_path / 'weights.nc'),
               discharge_files=[str(tmp_path / 'q.nc')], channel_state_init_file=c['state'], var_x='lon', var_y='lat', log=False)
    full, sub = Capture(), Capture()
    r0 = rr.RapidMuskingum(**cfg).set_write_discharges(full).route()
    r1 = rr.RapidMuskingum(**cfg).set_output_rivers(pick).set_write_discharges(sub).route()
    cols = [int(np.flatnonzero(ids == p)[0]) for p in pick]
    assert sub.calls[0][1].shape == (8, 4) and np.array_equal(sub.calls[0][1], full.calls[0][1][:, cols])
    assert np.array_equal(r1.channel_state, r0.channel_state)

    class Seam(rr.RapidMuskingum):                      # the reference's sequence on fp64 host arrays
        seen = None

        def _router(self, qlateral):
            Seam.seen = qlateral.shape
            q_t, arr = super()._router(qlateral)
            assert arr.shape == (8, 400) and arr.dtype == np.float64       # the seam always sees all segments in fp64
            return q_t, arr
    s2 = Capture()
    Seam(**cfg).set_output_rivers(pick).set_write_discharges(s2).route()
    assert Seam.seen == (8, 400) and np.array_equal(s2.calls[0][1], sub.calls[0][1])
    rr.RapidMuskingum(**cfg).set_output_rivers(pick).route()               # default writer stores the subset ids
    with ncio.open_nc(tmp_path / 'q.nc') as ds:
        assert np.array_equal(ncio.read_array(ds.variables['river_id']), pick.astype(np.int32))
        assert ncio.read_array(ds.variables['Q']).shape == (8, 4)
    with pytest.raises(ValueError, match='ids not in the params file'):
        rr.RapidMuskingum(**cfg).set_output_rivers([int(ids.max()) + 5]).route()


def test_qlateral_files_two_files_chain_state(tmp_path, host_only):
    from river_route_b200.runoff import QlateralDataset
    from river_route_b200 import synth
    n, T = 300, 6
    down = synth.forest(n, 2, seed=3, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 3)
    ids = np.arange(n, dtype=np.int64) + 1
    params = str(tmp_path / 'p.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(params)
    files, laterals = [], []
    for f in range(2):
        ql = synth.lateral_volumes(T, n, 30 + f)
        t = (np.datetime64('2021-05-01') + (np.arange(T) + f * T) * np.timedelta64(3, 'h')).astype('datetime64[s]')
        QlateralDataset(ql, ids, t, 'm3').to_netcdf(str(tmp_path / f'ql_{f}.nc'))
        files.append(str(tmp_path / f'ql_{f}.nc'))
        laterals.append(ql)
    r = rr.RapidMuskingum(params_file=params, qlateral_files=files, discharge_dir=str(tmp_path), dt_routing=3600,
                          log=False).route()                                # no state file: zero initial state + warning
    a = network_arrays(down, k, x, 3600, 10800)
    q = np.zeros(n)
    for f, ql in enumerate(laterals):
        ref = np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 3)
        with ncio.open_nc(tmp_path / f'discharge_ql_{f}.nc') as ds:
            assert np.array_equal(ncio.read_array(ds.variables['Q']), ref.astype(np.float32))
    assert np.array_equal(r.channel_state, q) and r.num_routing_steps_per_runoff == 3
    assert parity_error(r.channel_state, q) == 0.0


def test_route_twice_and_configs_instance(tmp_path, host_only):
    """A second route() on the same instance works (the reference keeps c1..c3 on self; here a rebuilt plan must get
    its coefficients again), and a Configs instance made with discharge_dir can be handed to a router."""
    from river_route_b200.runoff import QlateralDataset
    from river_route_b200 import synth
    n, T = 200, 5
    down = synth.forest(n, 2, seed=5, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 5)
    ids = np.arange(n, dtype=np.int64) + 1
    params = str(tmp_path / 'p.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(params)
    ql = synth.lateral_volumes(T, n, 41)
    t = (np.datetime64('2021-05-01') + np.arange(T) * np.timedelta64(1, 'h')).astype('datetime64[s]')
    QlateralDataset(ql, ids, t, 'm3').to_netcdf(str(tmp_path / 'ql.nc'))
    cfg = rr.Configs(params_file=params, qlateral_files=[str(tmp_path / 'ql.nc')], discharge_dir=str(tmp_path), log=False)
    assert cfg.discharge_files == [str(tmp_path / 'discharge_ql.nc')]
    r = rr.RapidMuskingum(cfg)                                              # used to raise 'not both'
    assert r.cfg is cfg
    r2 = rr.RapidMuskingum(cfg, dt_routing=1800)                            # kwargs still override a Configs instance
    assert r2.cfg.dt_routing == 1800 and r2.cfg.discharge_files == cfg.discharge_files
    a = network_arrays(down, k, x, 3600, 3600)
    q = np.zeros(n)
    cap = Capture()
    r.set_write_discharges(cap)
    for call in range(2):                                                   # state chains from the first call into the second
        r.route()
        plan_first = r.plan if call == 0 else plan_first
        ref = np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 1)
        assert np.array_equal(cap.calls[call][1], ref.astype(np.float32))
        assert hasattr(r.plan, '_fake_arrays')                              # set_coefficients reached the plan in use
    assert r.plan is plan_first                                             # unchanged network: the plan is reused
    # a changed params file (the first basin only) gives a new plan, and the new plan gets coefficients
    m = int(np.flatnonzero(down < 0)[0]) + 1
    assert 0 < m < n and not np.any(down[:m] >= m)
    pd.DataFrame({'river_id': ids[:m], 'downstream_river_id': np.where(down[:m] >= 0, ids[np.where(down[:m] >= 0, down[:m], 0)], -1),
                  'k': k[:m], 'x': x[:m]}).to_parquet(params)
    QlateralDataset(ql[:, :m], ids[:m], t, 'm3').to_netcdf(str(tmp_path / 'ql.nc'))
    del r.channel_state
    r.route()
    assert r.plan is not plan_first and hasattr(r.plan, '_fake_arrays') and r.n == m
    am = network_arrays(down[:m], k[:m], x[:m], 3600, 3600)
    qm, refm = np.zeros(m), np.zeros((T, m))
    oracle.rapid_route(am['indptr'], am['indices'], am['lhs_off'], am['c2'], am['c3'], am['c4_dt'], qm, ql[:, :m].copy(), refm, 1)
    assert np.array_equal(cap.calls[2][1], refm.astype(np.float32))


@pytest.mark.parametrize('router,f32,k,injected', [('rapid', False, 1, False), ('rapid', True, 2, False), ('rapid', True, 1, True),
                                                   ('unit', False, 2, False)])
def test_qlateral_files_streamed_in_slabs(tmp_path, host_only, monkeypatch, router, f32, k, injected):
    """Slab streaming (16-row slabs over 40-row files: 16 + 16 + 8) writes exactly what one whole-file call computes:
    float32 qlateral variables are routed as stored, dt_discharge resampling falls on slab boundaries, the channel
    state and the UH carry-over chain through slabs and files, and an injected writer still gets one array per file."""
    from river_route_b200.runoff import QlateralDataset
    from river_route_b200 import synth
    monkeypatch.setenv('RR_ROUTER_SLAB_ROWS', '16')
    n, T = 260, 40
    down = synth.forest(n, 3, seed=8, depth_bias=0.6)
    kk, x = synth.muskingum_params(n, 8)
    ids = np.arange(n, dtype=np.int64) + 100
    params = str(tmp_path / 'p.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': kk, 'x': x}).to_parquet(params)
    files, laterals = [], []
    for f in range(2):
        ql = synth.lateral_volumes(T, n, 50 + f) * (1.0 if router == 'rapid' else 1e-7)
        if f32:
            ql = ql.astype(np.float32)
        t = (np.datetime64('2022-01-01') + (np.arange(T) + f * T) * np.timedelta64(1, 'h')).astype('datetime64[s]')
        ds = QlateralDataset(ql, ids, t, 'm3')
        ds.to_netcdf(str(tmp_path / f'ql_{f}.nc'))
        if f32:                                          # store the variable as float32
            with ncio.open_nc(str(tmp_path / f'ql_{f}.nc'), 'w') as nc:
                nc.createDimension('time', T)
                nc.createDimension('river_id', n)
                tv = nc.createVariable('time', 'f8', ('time',))
                tv.units = 'seconds since 2022-01-01 00:00:00'
                tv[:] = (np.arange(T) + f * T) * 3600.0
                nc.createVariable('river_id', 'i4', ('river_id',))[:] = ids.astype(np.int32)
                nc.createVariable('qlateral', 'f4', ('time', 'river_id'))[:] = ql
        files.append(str(tmp_path / f'ql_{f}.nc'))
        laterals.append(ql.astype(np.float64))
    cfg = dict(params_file=params, qlateral_files=files, discharge_dir=str(tmp_path), dt_discharge=3600 * k, log=False)
    cap = Capture()
    if router == 'unit':
        rng = np.random.default_rng(2)
        ker = rng.uniform(0, 1, (5, n)) * (rng.random((5, n)) < 0.7)
        scipy.sparse.save_npz(str(tmp_path / 'uh.npz'), scipy.sparse.csr_matrix(ker))
        r = rr.UnitMuskingum(uh_kernel_file=str(tmp_path / 'uh.npz'), **cfg)
    else:
        r = rr.RapidMuskingum(**cfg)
    if injected:
        r.set_write_discharges(cap)
    r.route()
    a = network_arrays(down, kk, x, 3600, 3600)
    q = np.zeros(n)
    st = np.zeros((5, n))
    for f, ql in enumerate(laterals):
        ref = np.zeros((T, n))
        if router == 'unit':
            conv = oracle.uh_convolve(ql, ker, st)
            sp = oracle.unit_split(down.astype(np.int64))
            inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
            c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
            q_ch = q[inner].copy()
            q_fu = q_ch.copy()
            oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                              q_ch, q_fu, conv, ref, 1)
            q[hw] = conv[-1][hw]
            q[inner] = q_fu
        else:
            oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 1)
        want = (ref.reshape(T // k, k, n).mean(axis=1) if k > 1 else ref).astype(np.float32)
        if injected:
            dates, got, q_file, routed = cap.calls[f]
            assert got.shape == want.shape and dates.shape[0] == T // k and routed == files[f]
        else:
            with ncio.open_nc(tmp_path / f'discharge_ql_{f}.nc') as ds:
                got = ncio.read_array(ds.variables['Q'])
                assert ncio.read_array(ds.variables['time']).shape[0] == T // k
        if router == 'unit':
            # slab-wise UH convolution re-associates nothing (chunk-invariant by construction); the oracle's FFT-free
            # direct sum is the same arithmetic
            assert parity_error(got, want) < 1e-6
        else:
            assert np.array_equal(got, want), (f, router, f32, k)
    assert parity_error(r.channel_state, q) < 1e-12


@pytest.mark.parametrize('f32,k,slab_rows,group_bytes', [(False, 1, 0, 0), (True, 2, 16, 0), (False, 1, 16, 16 * 180 * 8 * 2)])
def test_ensemble_mode_batches_members(tmp_path, host_only, monkeypatch, f32, k, slab_rows, group_bytes):
    """runoff_processing_mode='ensemble' with qlateral files: members are routed by batched device calls (one per time
    slab and member group), every member from the same state; each member's file equals a single-member run and the
    final state is the member mean in file order (TransformMuskingum.py:121-126, :145-146)."""
    from river_route_b200.runoff import QlateralDataset
    from river_route_b200 import synth
    if slab_rows:
        monkeypatch.setenv('RR_ROUTER_SLAB_ROWS', str(slab_rows))
    if group_bytes:
        monkeypatch.setenv('RR_ROUTER_SLAB_BYTES', str(group_bytes))      # two members per group
    n, T, M = 180, 40, 5
    down = synth.forest(n, 3, seed=12, depth_bias=0.6)
    kk, x = synth.muskingum_params(n, 12)
    ids = np.arange(n, dtype=np.int64) + 7
    params = str(tmp_path / 'p.parquet')
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': kk, 'x': x}).to_parquet(params)
    q0 = np.random.default_rng(5).uniform(0, 25, n)
    pd.DataFrame({'Q': q0}).to_parquet(tmp_path / 'q0.parquet')
    t = (np.datetime64('2023-03-01') + np.arange(T) * np.timedelta64(1, 'h')).astype('datetime64[s]')
    files, laterals = [], []
    base = synth.lateral_volumes(T, n, 60)
    for m in range(M):
        ql = base * np.random.default_rng(70 + m).lognormal(0, 0.3)
        if f32:
            ql = ql.astype(np.float32)
        path = str(tmp_path / f'member_{m}.nc')
        with ncio.open_nc(path, 'w') as nc:
            nc.createDimension('time', T)
            nc.createDimension('river_id', n)
            tv = nc.createVariable('time', 'f8', ('time',))
            tv.units = 'seconds since 2023-03-01 00:00:00'
            tv[:] = np.arange(T) * 3600.0
            nc.createVariable('river_id', 'i4', ('river_id',))[:] = ids.astype(np.int32)
            nc.createVariable('qlateral', 'f4' if f32 else 'f8', ('time', 'river_id'))[:] = ql
        files.append(path)
        laterals.append(ql.astype(np.float64))
    calls = []
    real = plan_mod.Plan.route_ensemble_host

    def spy(self, mode, q_init, lats, outs, substeps, resample=1, q_final=None):
        calls.append((len(outs), outs[0].shape[0] * resample, q_init.ndim))
        return real(self, mode, q_init, lats, outs, substeps, resample=resample, q_final=q_final)
    monkeypatch.setattr(plan_mod.Plan, 'route_ensemble_host', spy)
    r = rr.RapidMuskingum(params_file=params, qlateral_files=files, discharge_dir=str(tmp_path), dt_discharge=3600 * k,
                          channel_state_init_file=str(tmp_path / 'q0.parquet'), runoff_processing_mode='ensemble',
                          channel_state_final_file=str(tmp_path / 'final.parquet'), log=False).route()
    assert calls and all(c[0] > 1 for c in calls if group_bytes == 0) and sum(c[0] for c in calls if c[2] == 1) == M
    if slab_rows:
        assert [c[1] for c in calls][:3] == [16, 16, 8]                   # slabs chain the members' own states
    if group_bytes:
        assert {c[0] for c in calls} == {2, 1}                            # groups of two, then the odd member
    a = network_arrays(down, kk, x, 3600, 3600)
    finals = []
    for m, ql in enumerate(laterals):
        q, ref = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 1)
        finals.append(q)
        want = (ref.reshape(T // k, k, n).mean(axis=1) if k > 1 else ref).astype(np.float32)
        with ncio.open_nc(tmp_path / f'discharge_member_{m}.nc') as ds:
            assert np.array_equal(ncio.read_array(ds.variables['Q']), want), m
    assert np.array_equal(r.channel_state, np.array(finals).mean(axis=0))
    assert np.array_equal(pd.read_parquet(tmp_path / 'final.parquet')['Q'].values, r.channel_state)
