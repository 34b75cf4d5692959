"""
Grid runoff -> catchment lateral inflow with the interface of river_route.runoff.runoff_to_qlateral
(river_route/runoff.py:218-378).  Table handling and file I/O are host Python (pandas / xarray) as in the
reference; the weight SpMM and its element-wise tail run on the GPU (rr_weights_transform_* of
include/rr_b200.h).  ``weights_to_qlateral`` is the array-level core and needs no netCDF stack.
"""
from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from .transforms import weights_transform

logger = logging.getLogger(__name__)

__all__ = ['runoff_to_qlateral', 'weights_to_qlateral', 'build_weight_csr']


def _conversion_factor(unit):
    if unit is None:
        logger.warning('No units attribute found. Assuming meters')
        return 1
    if unit in ('m', 'meters', 'kg m-2'):
        return 1
    if unit in ('mm', 'millimeters'):
        return .001
    raise ValueError(f'Unknown units: {unit}')


def build_weight_csr(river_id, x_index, y_index, proportion, area_sqm, conversion_factor=1):
    """
    Weight table rows -> (CSR indptr, indices, data, unique (x, y) cells, river ids in table order, catchment area).
    Cells are numbered in first-appearance order and rivers in first-appearance order (runoff.py:257-265, 283-290);
    duplicate (river, cell) rows are summed and columns sorted ascending, as scipy's COO -> CSR does (:292-295).
    """
    river_id = np.asarray(river_id)
    x_index = np.asarray(x_index).astype(np.int64)
    y_index = np.asarray(y_index).astype(np.int64)
    cell_key = pd.MultiIndex.from_arrays([x_index, y_index])
    point_idx, uniq_cells = pd.factorize(cell_key)          # first-appearance numbering == drop_duplicates order
    river_idx, river_ids_ordered = pd.factorize(river_id)
    n_riv, n_pts = len(river_ids_ordered), len(uniq_cells)
    vals = np.asarray(proportion, dtype=np.float64) * conversion_factor
    order = np.lexsort((np.arange(river_idx.shape[0]), point_idx, river_idx))
    r, p, v = river_idx[order], point_idx[order], vals[order]
    first = np.ones(r.shape[0], dtype=bool)
    first[1:] = (r[1:] != r[:-1]) | (p[1:] != p[:-1])
    starts = np.flatnonzero(first)
    data = np.add.reduceat(v, starts) if starts.size else np.zeros(0)
    indptr = np.zeros(n_riv + 1, dtype=np.int32)
    np.add.at(indptr, r[first] + 1, 1)
    np.cumsum(indptr, out=indptr)
    area = pd.Series(np.asarray(area_sqm, dtype=np.float64)).groupby(river_idx).sum().reindex(np.arange(n_riv)).to_numpy()
    cells_x = np.asarray(uniq_cells.get_level_values(0), dtype=np.int64)
    cells_y = np.asarray(uniq_cells.get_level_values(1), dtype=np.int64)
    return indptr, p[first].astype(np.int32), data, cells_x, cells_y, np.asarray(river_ids_ordered), area


def weights_to_qlateral(weight_table: dict, runoff_raw_or_grid: np.ndarray, *, runoff_depth_unit='m',
                        cumulative=False, force_positive_runoff=False, as_volumes=False):
    """
    Array-level runoff_to_qlateral.  ``weight_table`` has the columns river_id, x_index, y_index, proportion,
    area_sqm; the runoff is either the full grid (T, ny, nx) or the already gathered (T, n_points) array in
    unique-cell order.  Returns (qlateral (T, n_rivers) fp64, river ids in table order).
    """
    indptr, indices, data, cx, cy, rivers, area = build_weight_csr(
        weight_table['river_id'], weight_table['x_index'], weight_table['y_index'], weight_table['proportion'],
        weight_table['area_sqm'], _conversion_factor(runoff_depth_unit))
    raw = runoff_raw_or_grid
    if raw.ndim == 3:
        raw = raw[:, cy, cx]                                  # pointwise gather (runoff.py:270-279)
    ql = weights_transform(indptr, indices, data, raw, cumulative=cumulative, force_positive=force_positive_runoff,
                           area=area if as_volumes else None)
    return ql, rivers


def runoff_to_qlateral(runoff_data, grid_weights_file, *, var_runoff='ro', var_x='lon', var_y='lat', var_t='time',
                       var_river_id='river_id', runoff_depth_unit=None, cumulative=False, force_positive_runoff=False,
                       force_uniform_timesteps=True, as_volumes=False):
    """File-level entry point with the reference's signature; returns an xarray.Dataset like the reference."""
    try:
        import xarray as xr
    except ImportError as e:  # pragma: no cover - depends on the host environment
        raise ImportError('xarray is required for runoff_to_qlateral on files; use weights_to_qlateral for arrays') from e
    with xr.open_dataset(grid_weights_file) as ds:
        wdf = ds[[var_river_id, 'x_index', 'y_index', 'proportion', 'area_sqm']].to_dataframe()
    with xr.open_mfdataset(runoff_data) as ds:
        unit = runoff_depth_unit or ds[var_runoff].attrs.get('units', 'm')
        indptr, indices, data, cx, cy, rivers, area = build_weight_csr(
            wdf[var_river_id].values, wdf['x_index'].values, wdf['y_index'].values, wdf['proportion'].values,
            wdf['area_sqm'].values, _conversion_factor(unit))
        raw = (ds[var_runoff].isel({var_x: xr.DataArray(cx, dims='points'), var_y: xr.DataArray(cy, dims='points')})
               .transpose(var_t, 'points').values)
        time_index = ds[var_t].to_numpy()
    uniform = np.all(np.diff(time_index) == time_index[1] - time_index[0]) if len(time_index) > 1 else True
    resample = (not uniform) and force_uniform_timesteps
    # the non-uniform-time resample (runoff.py:316-329, rare) works on incremental depths before NaN / area handling
    ql = weights_transform(indptr, indices, data, raw, cumulative=cumulative, force_positive=force_positive_runoff,
                           area=None if resample else (area if as_volumes else None))
    if resample:
        timestep = int((time_index[1] - time_index[0]) / np.timedelta64(1, 's'))
        logger.warning(f'Time steps are not uniform, resampling to the first timestep: {timestep} seconds')
        df = pd.DataFrame(ql, index=time_index, columns=rivers).cumsum().resample(rule=f'{timestep}s').interpolate(method='linear')
        ql = np.vstack([df.values[0, :], np.diff(df.values, axis=0)])
        time_index = df.index.values
        ql[np.isnan(ql)] = 0.0
        if as_volumes:
            ql *= area[np.newaxis, :]
    units = 'm3' if as_volumes else 'm'
    return xr.Dataset(
        {'qlateral': xr.DataArray(ql, dims=('time', 'river_id'), attrs={'units': units})},
        coords={'river_id': xr.DataArray(np.asarray(rivers).astype(np.int64), dims=('river_id',)),
                'time': xr.DataArray(time_index, dims=('time',))})
