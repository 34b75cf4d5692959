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

__all__ = ['runoff_to_qlateral', 'weights_to_qlateral', 'build_weight_csr', 'read_weight_table', 'gather_grid_runoff',
           'grid_runoff_unit', 'grid_layout', 'QlateralDataset']


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
    # first-appearance numbering == drop_duplicates order; (x, y) packed into one integer key (a hash factorize of
    # int64 keys and a stable integer sort are ~10x faster than the tuple / lexsort route at 10^7-10^8 table rows)
    if x_index.size and (x_index.min() < 0 or y_index.min() < 0):
        raise ValueError('x_index / y_index must be non-negative grid indices')
    y_span = int(y_index.max()) + 1 if y_index.size else 1
    point_idx, uniq_keys = pd.factorize(x_index * y_span + y_index)
    river_idx, river_ids_ordered = pd.factorize(river_id)
    n_riv, n_pts = len(river_ids_ordered), len(uniq_keys)
    vals = np.asarray(proportion, dtype=np.float64) * conversion_factor
    order = np.argsort(river_idx.astype(np.int64) * max(n_pts, 1) + point_idx, kind='stable')   # by (river, cell), ties in file order
    r, p, v = river_idx[order], point_idx[order], vals[order]
    first = np.ones(r.shape[0], dtype=bool)
    first[1:] = (r[1:] != r[:-1]) | (p[1:] != p[:-1])
    starts = np.flatnonzero(first)
    data = np.add.reduceat(v, starts) if starts.size else np.zeros(0)
    indptr = np.zeros(n_riv + 1, dtype=np.int32)
    np.cumsum(np.bincount(r[first], minlength=n_riv), out=indptr[1:])
    area = pd.Series(np.asarray(area_sqm, dtype=np.float64)).groupby(river_idx).sum().reindex(np.arange(n_riv)).to_numpy()
    uniq_keys = np.asarray(uniq_keys, dtype=np.int64)
    cells_x, cells_y = uniq_keys // y_span, uniq_keys % y_span
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


def read_weight_table(grid_weights_file, var_river_id='river_id') -> dict:
    """Columns of the weight table netCDF (docs/references/io-file-schema.md:76-84) in file (index) order."""
    from . import ncio
    with ncio.open_nc(grid_weights_file) as ds:
        cols = {}
        for key, name in (('river_id', var_river_id), ('x_index', 'x_index'), ('y_index', 'y_index'),
                          ('proportion', 'proportion'), ('area_sqm', 'area_sqm')):
            if name not in ds.variables:
                raise KeyError(f'{name} not found in weight table {grid_weights_file}')
            cols[key] = ncio.read_array(ds.variables[name])
    return cols


def _as_list(runoff_data):
    import os
    return [runoff_data] if isinstance(runoff_data, (str, os.PathLike)) else list(runoff_data)


def grid_runoff_unit(runoff_data, var_runoff='ro'):
    """``units`` attribute of the runoff variable; 'm' when the file has none (runoff.py:268)."""
    from . import ncio
    with ncio.open_nc(_as_list(runoff_data)[0]) as ds:
        return ncio.attrs_of(ds.variables[var_runoff]).get('units', 'm')


def grid_layout(runoff_data, *, var_runoff='ro', var_x='lon', var_y='lat', var_t='time'):
    """(dimension names of the runoff variable in file order, ny, nx) of the first runoff file."""
    from . import ncio
    with ncio.open_nc(_as_list(runoff_data)[0]) as ds:
        var = ds.variables[var_runoff]
        dims = tuple(var.dimensions)
        if sorted(dims) != sorted((var_t, var_y, var_x)):
            raise ValueError(f'{var_runoff} must have dimensions ({var_t}, {var_y}, {var_x}), found {dims}')
        return dims, int(var.shape[dims.index(var_y)]), int(var.shape[dims.index(var_x)])


def gather_grid_runoff(runoff_data, cells_x, cells_y, *, var_runoff='ro', var_x='lon', var_y='lat', var_t='time',
                       slab_rows=64, flat=False):
    """
    The pointwise gather of runoff.py:267-280: (time axis as datetime64[s], (T, n_points) runoff in the file's dtype)
    for the unique cells (x_index, y_index) of the weight table.  Files are concatenated along time like
    ``xr.open_mfdataset`` does for consecutive files; the grid is read in slabs of time steps.
    ``flat=True`` skips the gather and returns the whole grid as (T, ny * nx) -- for weight tables whose column
    indices are flat cell ids ``y * nx + x``, i.e. the gather happens inside the device SpMM (files stored as
    (time, y, x) only).
    """
    from . import ncio
    dates, parts = [], []
    for path in _as_list(runoff_data):
        with ncio.open_nc(path) as ds:
            var = ds.variables[var_runoff]
            dims = tuple(var.dimensions)
            if sorted(dims) != sorted((var_t, var_y, var_x)):
                raise ValueError(f'{var_runoff} must have dimensions ({var_t}, {var_y}, {var_x}), found {dims}')
            at, ay, ax = dims.index(var_t), dims.index(var_y), dims.index(var_x)
            tv = ds.variables[var_t]
            dates.append(ncio.decode_time(ncio.read_array(tv), ncio.attrs_of(tv).get('units', '')))
            T = var.shape[at]
            for t0 in range(0, T, slab_rows):
                sel = [slice(None)] * 3
                sel[at] = slice(t0, min(T, t0 + slab_rows))
                slab = np.moveaxis(np.asarray(var[tuple(sel)]), (at, ay, ax), (0, 1, 2))
                if isinstance(slab, np.ma.MaskedArray):
                    slab = slab.filled(np.nan)
                if flat:
                    if (at, ay, ax) != (0, 1, 2):
                        raise ValueError('flat grids need the runoff stored as (time, y, x)')
                    g = slab.reshape(slab.shape[0], -1)
                else:
                    g = slab[:, cells_y, cells_x]
                parts.append(np.ascontiguousarray(g, dtype=g.dtype.newbyteorder('=')))
    return np.concatenate(dates), np.concatenate(parts, axis=0)


class QlateralDataset(dict):
    """What ``runoff_to_qlateral`` returns when xarray is not installed: ``ds['qlateral'].values``,
    ``ds['time'].values``, ``ds['river_id'].values`` and ``to_netcdf`` with the reference's layout
    (docs/references/io-file-schema.md:52-55)."""

    class _Var:
        def __init__(self, values, attrs=None):
            self.values, self.attrs = values, dict(attrs or {})

        def to_numpy(self):
            return self.values

    def __init__(self, ql, rivers, time_index, units):
        super().__init__(qlateral=self._Var(ql, {'units': units}), river_id=self._Var(np.asarray(rivers).astype(np.int64)),
                         time=self._Var(np.asarray(time_index)))

    def to_netcdf(self, path):
        from . import ncio
        t = self['time'].values.astype('datetime64[s]')
        with ncio.open_nc(path, 'w') as ds:
            ds.createDimension('time', t.shape[0])
            ds.createDimension('river_id', self['river_id'].values.shape[0])
            tv = ds.createVariable('time', 'f8', ('time',))
            tv.units = f'seconds since {pd.Timestamp(t[0]).strftime("%Y-%m-%d %H:%M:%S")}'
            tv[:] = (t - t[0]).astype('timedelta64[s]').astype(np.float64)
            rv = ds.createVariable('river_id', 'i4', ('river_id',))
            rv[:] = self['river_id'].values.astype(np.int32)
            qv = ds.createVariable('qlateral', 'f8', ('time', 'river_id'))
            qv.units = self['qlateral'].attrs.get('units', 'm')
            qv[:] = self['qlateral'].values


def resample_irregular(ql, time_index, rivers, area=None):
    """
    The rare irregular-time-axis branch of ``runoff_to_qlateral`` (runoff.py:316-337), kept on the host with pandas as
    the reference has it: incremental depths -> cumulative -> resample to the first time step -> linear interpolation
    -> incremental; only then NaN -> 0 and the optional multiplication by catchment area.  ``ql`` must still carry
    its NaNs (``weights_transform(..., keep_nan=True)``): the reference lets them run through cumsum / interpolate.
    """
    timestep = int((time_index[1] - time_index[0]) / np.timedelta64(1, 's'))
    logger.warning(f'Time steps are not uniform, resampling to the first timestep: {timestep} seconds')
    df = pd.DataFrame(ql, index=time_index, columns=rivers).cumsum().resample(rule=f'{timestep}s').interpolate(method='linear')
    out = np.vstack([df.values[0, :], np.diff(df.values, axis=0)])
    out[np.isnan(out)] = 0.0
    if area is not None:
        out *= np.asarray(area)[np.newaxis, :]
    return out, df.index.values


def runoff_to_qlateral(runoff_data, grid_weights_file, *, var_runoff='ro', var_x='lon', var_y='lat', var_t='time',
                       var_river_id='river_id', runoff_depth_unit=None, cumulative=False, force_positive_runoff=False,
                       force_uniform_timesteps=True, as_volumes=False):
    """File-level entry point with the reference's signature (runoff.py:218-378).  Returns an ``xarray.Dataset``
    like the reference when xarray is installed, else a :class:`QlateralDataset` with the same variables."""
    table = read_weight_table(grid_weights_file, var_river_id)
    unit = runoff_depth_unit or grid_runoff_unit(runoff_data, var_runoff)
    indptr, indices, data, cx, cy, rivers, area = build_weight_csr(
        table['river_id'], table['x_index'], table['y_index'], table['proportion'], table['area_sqm'],
        _conversion_factor(unit))
    time_index, raw = gather_grid_runoff(runoff_data, cx, cy, var_runoff=var_runoff, var_x=var_x, var_y=var_y, var_t=var_t)
    uniform = np.all(np.diff(time_index) == time_index[1] - time_index[0]) if len(time_index) > 1 else True
    resample = (not uniform) and force_uniform_timesteps
    # the non-uniform-time resample (runoff.py:316-329, rare) works on incremental depths before NaN / area handling
    ql = weights_transform(indptr, indices, data, raw, cumulative=cumulative, force_positive=force_positive_runoff,
                           area=None if resample else (area if as_volumes else None), keep_nan=resample)
    if resample:
        ql, time_index = resample_irregular(ql, time_index, rivers, area if as_volumes else None)
    units = 'm3' if as_volumes else 'm'
    try:
        import xarray as xr
    except ImportError:
        return QlateralDataset(ql, rivers, time_index, units)
    return xr.Dataset(
        {'qlateral': xr.DataArray(ql, dims=('time', 'river_id'), attrs={'units': units})},
        coords={'river_id': xr.DataArray(np.asarray(rivers).astype(np.int64), dims=('river_id',)),
                'time': xr.DataArray(time_index, dims=('time',))})
