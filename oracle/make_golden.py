"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code
(/root/reference, river-route v2.0.1: its numba kernels, router classes, UnitHydrograph, scipy
calls) on small seeded inputs.  Run in the build container only:

    python oracle/make_golden.py

/root/reference does not exist on the GPU box, so the outputs are committed as small .npz files and
this script is kept next to them as their provenance.  xarray / netCDF4 / geopandas / shapely are
not installed here; they are only needed by the reference at import time for annotations, so empty
stub modules are registered.  The routers are driven exactly as their own
route() does (Muskingum.py:199-227), minus file I/O: params come from a real parquet file, lateral
inflow is injected by overriding ``_qlateral_generator`` and the fp64 arrays are taken from
``_router`` before the float32 cast.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(tempfile.gettempdir(), 'rr_numba_cache'))
sys.dont_write_bytecode = True

import numpy as np
import pandas as pd
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub('xarray', Dataset=type('Dataset', (), {}), DataArray=type('DataArray', (), {}))
    stub('netCDF4')
    stub('geopandas', GeoDataFrame=type('GeoDataFrame', (), {}))
    stub('shapely')
    stub('shapely.geometry', Point=object, MultiPoint=object, box=object)
    stub('shapely.ops', voronoi_diagram=object)
    sys.path.insert(0, REF)
    import river_route  # noqa
    return river_route


def random_forest(n, n_basins, rng, p_two=0.7, depth_bias=0.6):
    """Small pure-Python forest generator (independent of the product's generator)."""
    sizes = np.maximum(1, np.floor(rng.dirichlet(np.ones(n_basins)) * (n - n_basins)).astype(int) + 1)
    sizes[np.argmax(sizes)] += n - sizes.sum()
    down = np.full(n, -1, dtype=np.int64)
    off = 0
    for m in sizes:
        parent = [-1]
        tips = [0]
        while len(parent) < m:
            pick = len(tips) - 1 if rng.random() < depth_bias else int(rng.integers(len(tips)))
            t = tips.pop(pick)
            for _ in range(min(2 if rng.random() < p_two else 1, m - len(parent))):
                parent.append(t)
                tips.append(len(parent) - 1)
        for g, pg in enumerate(parent):
            down[off + m - 1 - g] = -1 if pg < 0 else off + m - 1 - pg
        off += m
    return down


def shuffle_topological(down, rng):
    """Random valid upstream-before-downstream relabelling."""
    import heapq
    n = down.shape[0]
    prio = rng.permutation(n)
    indeg = np.bincount(down[down >= 0], minlength=n)
    heap = [(int(prio[i]), int(i)) for i in np.flatnonzero(indeg == 0)]
    heapq.heapify(heap)
    order = []
    while heap:
        _, i = heapq.heappop(heap)
        order.append(i)
        d = int(down[i])
        if d >= 0:
            indeg[d] -= 1
            if indeg[d] == 0:
                heapq.heappush(heap, (int(prio[d]), d))
    order = np.array(order)
    new_of_old = np.empty(n, dtype=np.int64)
    new_of_old[order] = np.arange(n)
    d_old = down[order]
    return np.where(d_old >= 0, new_of_old[np.where(d_old >= 0, d_old, 0)], -1)


def write_params(path, down, k, x, rng):
    """params parquet with non-trivial river ids (io-file-schema.md:15-20)."""
    n = down.shape[0]
    ids = (rng.permutation(n) + 1) * 7 + 100000
    pd.DataFrame({
        'river_id': ids.astype(np.int64),
        'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1).astype(np.int64),
        'k': k, 'x': x,
    }).to_parquet(path)
    return ids


def lateral(T, n, rng, scale):
    a = rng.gamma(0.3, scale, size=(T, n))
    a[rng.random((T, n)) < 0.5] = 0.0
    return a


def main():
    rr = import_reference()
    from river_route.routers import _numba_kernels as nk
    from river_route.tools import adjacency_matrix
    from river_route.uhkernels import SCSTriangular, UnitHydrograph
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(20260218)

    cases = [
        # name,        n,   basins, T,  dt_runoff, dt_routing, shuffle
        ('small',      137,  2,     18, 10800,     10800,      False),
        ('substeps',   301,  3,     12, 10800,     900,        False),
        ('shuffled',   260,  4,     15, 3600,      1800,       True),
        ('chain',      97,   1,     14, 3600,      3600,       False),
    ]
    for name, n, nbas, T, dt_runoff, dt_routing, shuf in cases:
        down = random_forest(n, nbas, rng, depth_bias=0.97 if name == 'chain' else 0.6)
        if shuf:
            down = shuffle_topological(down, rng)
        k = rng.uniform(1800.0, 20000.0, n)
        x = rng.uniform(0.05, 0.4, n)
        params = os.path.join(tmp, f'{name}.parquet')
        ids = write_params(params, down, k, x, rng)
        q0 = rng.uniform(0.0, 40.0, n)
        state = os.path.join(tmp, f'{name}_state.parquet')
        pd.DataFrame({'Q': q0}).to_parquet(state)
        dates = (np.datetime64('2020-01-01T00:00:00') + np.arange(T) * np.timedelta64(dt_runoff, 's')).astype('datetime64[s]')

        # ---------------- RapidMuskingum (RapidMuskingum.py:19-33 -> rapid_route) ----------------
        ql = lateral(T, n, rng, 5.0e4)
        r = rr.RapidMuskingum(params_file=params, qlateral_files=[params], discharge_files=[os.path.join(tmp, 'o.nc')],
                              channel_state_init_file=state, dt_routing=dt_routing, log=False)
        r._set_network_dependent_vectors()
        r._read_initial_state()
        r._set_network_and_time_dependent_vectors(dates)
        q_rapid, out_rapid = r._router(ql)
        golden = dict(
            down=down, river_ids=ids, k=k, x=x, q0=q0, dt_runoff=dt_runoff, dt_routing=dt_routing,
            csc_indptr=r._csc_indptr, csc_indices=r._csc_indices, lhs_off=r._lhs_off_data,
            c1=r.c1, c2=r.c2, c3=r.c3, c4_dt=r.c4 / r.dt_runoff, substeps=r.num_routing_steps_per_runoff,
            ql=ql, rapid_out=out_rapid, rapid_q=q_rapid,
        )
        A = adjacency_matrix(ids, np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1))
        assert np.array_equal(A.indptr, r._csc_indptr) and np.array_equal(A.indices, r._csc_indices)

        # ---------------- Muskingum channel only (Muskingum.py:262-290 -> muskingum_route) -------
        nrpo = 3
        n_out = 10
        m = rr.Muskingum(params_file=params, discharge_files=[os.path.join(tmp, 'o.nc')],
                         channel_state_init_file=state, dt_routing=dt_routing, dt_total=dt_routing * nrpo * n_out,
                         dt_discharge=dt_routing * nrpo, log=False)
        m._set_network_dependent_vectors()
        m._read_initial_state()
        m._set_muskingum_coefficients(dt_routing)
        out_m = m._router(n_out, nrpo)
        golden.update(musk_out=out_m, musk_q=m.channel_state, musk_nrpo=nrpo, musk_nout=n_out)

        # ---------------- UnitMuskingum (UnitMuskingum.py:31-98 -> convolve + unit_route) --------
        area = rng.uniform(1e5, 5e8, n)
        uh_builder = SCSTriangular(tc=5.0 * k, area=area, tr=float(dt_runoff))   # as tests/conftest.py:57-64
        kfile = os.path.join(tmp, f'{name}_uh.npz')
        uh_builder.save(kfile)
        depths = lateral(T, n, rng, 2.0e-3)
        u = rr.UnitMuskingum(params_file=params, qlateral_files=[params], discharge_files=[os.path.join(tmp, 'o.nc')],
                             channel_state_init_file=state, uh_kernel_file=kfile, dt_routing=dt_routing, log=False)
        u._set_network_dependent_vectors()
        u._read_initial_state()
        u._hook_before_route()
        u._set_network_and_time_dependent_vectors(dates)
        uh_state0 = rng.uniform(0.0, 5.0, u._uh.kernel.shape)
        uh_state0[-1] = 0.0
        u._uh.state = uh_state0.copy()
        # the convolved lateral the router hands to unit_route, recorded separately
        uh_probe = UnitHydrograph(kfile)
        uh_probe.state = uh_state0.copy()
        conv = np.ascontiguousarray(uh_probe.convolve(depths))
        q_unit, out_unit = u._router(depths)
        golden.update(
            uh_kernel=u._uh.kernel, uh_state0=uh_state0, uh_state1=u._uh.state.copy(), depths=depths, conv=conv,
            unit_out=out_unit, unit_q=q_unit, hw_idx=u.hw_idx, inner_idx=u.inner_idx,
            a_inner_indptr=u._a_inner_indptr, a_inner_indices=u._a_inner_indices,
            a_hw_indptr=u._a_hw_indptr, a_hw_indices=u._a_hw_indices,
        )
        # second file through the same router (sequential mode): state and UH carry-over chain
        u.channel_state = q_unit
        depths2 = lateral(T, n, rng, 2.0e-3)
        q_unit2, out_unit2 = u._router(depths2)
        golden.update(depths2=depths2, unit_out2=out_unit2, unit_q2=q_unit2, uh_state2=u._uh.state.copy())
        np.savez_compressed(os.path.join(OUT, f'route_{name}.npz'), **golden)
        print('wrote', name, {k_: np.asarray(v).shape for k_, v in golden.items() if np.asarray(v).ndim == 2})

    # ---------------- UnitHydrograph (UnitHydrograph.py:64-107) ------------------------------------
    # the reference's own data-free cases: tests/test_uhkernels.py:52-99
    np.random.seed(123)
    kernel = np.random.rand(3, 4)
    lat_a = np.random.rand(10, 4)
    kfile = os.path.join(tmp, 'k.npz')
    scipy.sparse.save_npz(kfile, scipy.sparse.csr_matrix(kernel))
    uh_full = UnitHydrograph(kfile)
    res_full = uh_full.convolve(lat_a)
    uh_inc = UnitHydrograph(kfile)
    res_inc = np.array([uh_inc.convolve_incrementally(lat_a[t]) for t in range(10)])
    # longer kernels, T < n_ks and carry-over across three calls
    kernel_b = rng.uniform(0.0, 1.0, (23, 57)) * (rng.random((23, 57)) < 0.6)
    kfile_b = os.path.join(tmp, 'kb.npz')
    scipy.sparse.save_npz(kfile_b, scipy.sparse.csr_matrix(kernel_b))
    uh_b = UnitHydrograph(kfile_b)
    uh_b_inc = UnitHydrograph(kfile_b)
    calls = [rng.gamma(0.3, 2e-3, (T_, 57)) for T_ in (40, 7, 1)]
    outs_b, states_b, outs_b_inc = [], [], []
    for c in calls:
        outs_b.append(uh_b.convolve(c).copy())
        states_b.append(uh_b.state.copy())
        outs_b_inc.append(np.array([uh_b_inc.convolve_incrementally(c[t]) for t in range(c.shape[0])]))
    np.savez_compressed(
        os.path.join(OUT, 'uh.npz'), kernel=kernel, lateral=lat_a, conv_full=res_full, conv_inc=res_inc,
        state_full=uh_full.state, state_inc=uh_inc.state,
        impulse_kernel=np.array([[1.0, 0.5], [0.5, 0.3], [0.0, 0.2]]),
        kernel_b=kernel_b, call0=calls[0], call1=calls[1], call2=calls[2],
        out0=outs_b[0], out1=outs_b[1], out2=outs_b[2], inc0=outs_b_inc[0], inc1=outs_b_inc[1], inc2=outs_b_inc[2],
        state0=states_b[0], state1=states_b[1], state2=states_b[2], state_inc_final=uh_b_inc.state)
    print('wrote uh')

    # ---------------- grid weights: the arithmetic of runoff.py:257-337 ------------------------------
    # xr.open_dataset / open_mfdataset cannot run here; everything between them is executed with the
    # reference's own pandas / scipy calls on in-memory frames (copied call sequence, not a restatement).
    n_riv, ncell_x, ncell_y, T = 83, 12, 9, 11
    rows = []
    for rid in range(n_riv):
        for _ in range(int(rng.integers(1, 7))):
            rows.append((1000 + rid, int(rng.integers(ncell_x)), int(rng.integers(ncell_y)), rng.random(), rng.uniform(1e5, 5e8)))
    rows.append(rows[5])  # an exact duplicate (river, cell) pair: scipy sums it
    weight_df = pd.DataFrame(rows, columns=['river_id', 'x_index', 'y_index', 'proportion', 'area_sqm'])
    weight_df['proportion'] /= weight_df.groupby('river_id')['proportion'].transform('sum')
    grid = rng.gamma(0.3, 2e-3, (T, ncell_y, ncell_x)).astype(np.float32)
    grid[rng.random(grid.shape) < 0.6] = 0.0
    grid[3, 2, 5] = np.nan
    var_river_id = 'river_id'
    unique_indexes = (weight_df[['x_index', 'y_index']].drop_duplicates().reset_index(drop=True).reset_index().astype(int))
    unique_sorted_rivers = weight_df[[var_river_id, ]].drop_duplicates().sort_index()
    runoff_raw = grid[:, unique_indexes['y_index'].values, unique_indexes['x_index'].values]   # pointwise isel, (T, points)
    point_idx = (weight_df[['x_index', 'y_index']].merge(unique_indexes, on=['x_index', 'y_index'], how='left')['index'].values)
    river_ids_ordered = unique_sorted_rivers[var_river_id].values
    river_id_to_row = pd.Series(np.arange(len(river_ids_ordered)), index=river_ids_ordered)
    river_idx = river_id_to_row.loc[weight_df[var_river_id].values].values
    golden_w = dict(river_id=weight_df['river_id'].values, x_index=weight_df['x_index'].values,
                    y_index=weight_df['y_index'].values, proportion=weight_df['proportion'].values,
                    area_sqm=weight_df['area_sqm'].values, grid=grid, runoff_raw=runoff_raw,
                    point_idx=point_idx, river_idx=river_idx, river_ids_ordered=river_ids_ordered)
    for conv_name, conversion_factor in (('m', 1), ('mm', .001)):
        weights = scipy.sparse.csr_matrix(
            (weight_df['proportion'].values * conversion_factor, (river_idx, point_idx)),
            shape=(len(river_ids_ordered), len(unique_indexes)))
        golden_w[f'csr_indptr_{conv_name}'] = weights.indptr
        golden_w[f'csr_indices_{conv_name}'] = weights.indices
        golden_w[f'csr_data_{conv_name}'] = weights.data
        catchment_area = (weight_df.groupby(var_river_id)['area_sqm'].sum().reindex(unique_sorted_rivers[var_river_id].values).to_numpy())
        golden_w['catchment_area'] = catchment_area
        for cumulative in (False, True):
            for as_volumes in (False, True):
                src = np.cumsum(runoff_raw.astype(np.float64), axis=0).astype(np.float32) if cumulative else runoff_raw
                qlateral = np.asarray(weights @ src.T).T
                if cumulative:
                    for i in range(qlateral.shape[0] - 1, 0, -1):
                        qlateral[i] -= qlateral[i - 1]
                mask = np.isnan(qlateral)
                if mask.any():
                    qlateral[mask] = 0.0
                if as_volumes:
                    qlateral *= catchment_area[np.newaxis, :]
                golden_w[f'ql_{conv_name}_cum{int(cumulative)}_vol{int(as_volumes)}'] = np.ascontiguousarray(qlateral)
                if cumulative:
                    golden_w['runoff_raw_cumulative'] = src
    np.savez_compressed(os.path.join(OUT, 'weights.npz'), **golden_w)
    print('wrote weights')

    # ---------------- irregular time axis: runoff.py:316-337 with NaN cells carried through the resample ----------
    # (round 2; a separate file and a separate generator so that the fixtures above stay byte-identical)
    from river_route.runoff import _cumulative_to_incremental, _incremental_to_cumulative
    rng_i = np.random.default_rng(20260218)
    Ti = 9
    hours = np.array([0, 3, 6, 9, 15, 21, 24, 30, 36])                    # 3-hourly, then 6-hourly
    time_index = (np.datetime64('2020-01-01T00:00:00') + hours.astype('timedelta64[h]')).astype('datetime64[ns]')
    grid_i = rng_i.gamma(0.3, 2e-3, (Ti, ncell_y, ncell_x)).astype(np.float32)
    grid_i[rng_i.random(grid_i.shape) < 0.5] = 0.0
    grid_i[4, 2, 5] = np.nan
    grid_i[0, 7, 1] = np.nan
    raw_i = grid_i[:, unique_indexes['y_index'].values, unique_indexes['x_index'].values]
    weights = scipy.sparse.csr_matrix((weight_df['proportion'].values * 1, (river_idx, point_idx)),
                                      shape=(len(river_ids_ordered), len(unique_indexes)))
    golden_i = dict(runoff_raw=raw_i, time_index=time_index.astype('datetime64[s]').astype(np.int64))
    for cumulative in (False, True):
        for as_volumes in (False, True):
            src = np.cumsum(raw_i.astype(np.float64), axis=0).astype(np.float32) if cumulative else raw_i
            qlateral = np.asarray(weights @ src.T).T
            if cumulative:
                for i in range(qlateral.shape[0] - 1, 0, -1):
                    qlateral[i] -= qlateral[i - 1]
            timestep = int((time_index[1] - time_index[0]) / np.timedelta64(1, 's'))
            df = pd.DataFrame(qlateral, index=time_index, columns=river_ids_ordered)
            df = (_incremental_to_cumulative(df).resample(rule=f'{timestep}s').interpolate(method='linear'))
            df = _cumulative_to_incremental(df)
            t_out = df.index.values
            # pandas 3 (copy-on-write) hands back a read-only view here, on which the reference's own in-place NaN
            # fill (runoff.py:331-333) raises; a writable copy gives what it computes with the pandas it was written for
            qlateral = df.to_numpy(dtype=np.float64).copy()
            mask = np.isnan(qlateral)
            if mask.any():
                qlateral[mask] = 0.0
            if as_volumes:
                qlateral *= catchment_area[np.newaxis, :]
            golden_i[f'ql_cum{int(cumulative)}_vol{int(as_volumes)}'] = np.ascontiguousarray(qlateral)
            golden_i['time_out'] = t_out.astype('datetime64[s]').astype(np.int64)
            if cumulative:
                golden_i['runoff_raw_cumulative'] = src
    np.savez_compressed(os.path.join(OUT, 'weights_irregular.npz'), **golden_i)
    print('wrote weights_irregular')

    # ---------------- tools.adjacency_matrix known answers (tools.py:75-109) -------------------------
    # 9-reach network of docs/references/math.md:70-80 and the rejection cases of tests/test_tools.py:48-60
    ids9 = np.arange(1, 10)
    ds9 = np.array([5, 5, 6, 6, 7, 7, 9, 9, -1])
    A9 = adjacency_matrix(ids9, ds9)
    errs = {}
    for key, (ri, di) in {'unsorted': (np.array([10, 20, 30]), np.array([20, -1, 10])),
                          'unknown': (np.array([10, 20]), np.array([-1, 999]))}.items():
        try:
            adjacency_matrix(ri, di)
            errs[key] = ''
        except ValueError as e:
            errs[key] = str(e)
    np.savez_compressed(os.path.join(OUT, 'tools.npz'), ids9=ids9, ds9=ds9, A9_dense=A9.toarray(), A9_indptr=A9.indptr,
                        A9_indices=A9.indices, err_unsorted=errs['unsorted'], err_unknown=errs['unknown'])
    print('wrote tools', errs)


if __name__ == '__main__':
    main()
