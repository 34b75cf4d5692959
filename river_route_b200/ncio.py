"""
netCDF access for the host side of the routers (file formats of docs/references/io-file-schema.md).

The reference reads with xarray and writes with netCDF4 (routers/TransformMuskingum.py:34-37, runoff.py:255-280,
routers/Muskingum.py:337-351).  This module talks to ``netCDF4`` when it is installed and otherwise to
``scipy.io.netcdf_file`` -- the classic (netCDF-3) format only -- so that the file-level paths also work, and are
tested, on machines without the HDF5 stack.  Both libraries expose the same small API used here
(``variables[name][:]``, ``.dimensions``, attributes, ``createDimension`` / ``createVariable``).
Pure host I/O: nothing in here computes.
"""
from __future__ import annotations

import re

import numpy as np

__all__ = ['open_nc', 'read_array', 'attrs_of', 'decode_time', 'backend']

_UNIT_SECONDS = {'second': 1, 'seconds': 1, 'sec': 1, 'secs': 1, 's': 1, 'minute': 60, 'minutes': 60, 'min': 60,
                 'mins': 60, 'hour': 3600, 'hours': 3600, 'hr': 3600, 'hrs': 3600, 'h': 3600, 'day': 86400, 'days': 86400,
                 'd': 86400}


def backend() -> str:
    try:
        import netCDF4  # noqa: F401
        return 'netCDF4'
    except ImportError:
        return 'scipy'


def open_nc(path, mode: str = 'r', mmap: bool = False):
    """Open (or create, ``mode='w'``) a netCDF file.  New files are NETCDF4 with netCDF4 installed (as the reference
    writes them, Muskingum.py:337), else 64-bit-offset classic files."""
    path = str(path)
    if backend() == 'netCDF4':
        import netCDF4
        return netCDF4.Dataset(path, mode=mode, format='NETCDF4') if mode == 'w' else netCDF4.Dataset(path, mode=mode)
    from scipy.io import netcdf_file
    if mode == 'w':
        return netcdf_file(path, mode='w', version=2)
    try:
        return netcdf_file(path, mode='r', mmap=bool(mmap), maskandscale=True)
    except TypeError as e:
        raise ImportError(f'{path} is not a classic netCDF-3 file; reading NETCDF4/HDF5 files needs the netCDF4 '
                          f'package') from e


def read_array(var) -> np.ndarray:
    """Values of a netCDF variable as a plain ndarray (masked entries -> NaN for floating point data)."""
    a = var[:]
    if isinstance(a, np.ma.MaskedArray):
        a = a.filled(np.nan) if a.dtype.kind == 'f' else a.filled()
    a = np.array(a, copy=True)
    if a.dtype.byteorder == '>':                          # classic files are big-endian on disk
        a = a.astype(a.dtype.newbyteorder('='))
    return a


def attrs_of(var) -> dict:
    if hasattr(var, 'ncattrs'):
        return {k: var.getncattr(k) for k in var.ncattrs()}
    out = {}
    for k, v in getattr(var, '_attributes', {}).items():
        out[k] = v.decode() if isinstance(v, bytes) else v
    return out


def decode_time(values, units: str) -> np.ndarray:
    """CF time axis ``<unit> since <date>`` -> datetime64[s] (standard calendar)."""
    m = re.match(r'\s*(\w+)\s+since\s+(.+?)\s*$', str(units))
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError(f'Unsupported time units: {units!r}')
    stamp = m.group(2).strip().replace('T', ' ')
    stamp = re.sub(r'\s*(UTC|Z|\+00:?00)$', '', stamp)
    stamp = re.sub(r'(\d{2}:\d{2}:\d{2})\.\d+$', r'\1', stamp)
    origin = np.datetime64(stamp.replace(' ', 'T'), 's')
    secs = np.rint(np.asarray(values, dtype=np.float64) * _UNIT_SECONDS[m.group(1).lower()]).astype(np.int64)
    return origin + secs.astype('timedelta64[s]')
