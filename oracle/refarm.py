"""
The reference's own CPU implementation of the routing path, timed: ``bench.py --impl reference`` and the
``cpu_baseline`` leg of ``bench.py``.  MEASUREMENT INFRASTRUCTURE, not product code -- nothing under
river_route_b200/ imports this module, and this module never imports river_route_b200 (so the reference arm's
process maps neither librr_b200.so nor any CUDA library).

What runs is river-route v2.0.1 itself, unmodified, from the copy ``oracle/ref_install.py`` put under oracle/_ref/:
``river_route.tools.adjacency_matrix`` (tools.py:75-109), ``Muskingum._set_muskingum_coefficients``
(routers/Muskingum.py:172-193) and ``RapidMuskingum._router`` (routers/RapidMuskingum.py:19-33), which allocates the
discharge array, copies the channel state and calls the numba kernel ``rapid_route`` (_numba_kernels.py:49-84).
The kernels are single-threaded by construction; the all-cores figure is obtained the way the reference
documents it -- independent watersheds in separate processes (docs/references/parallelism.md:67-112).

The synthetic network comes from oracle/_build/librr_hostutil.so, the stand-alone build of the very source file
(river_route_b200/csrc/rr_hostutil.cpp) the product library compiles, so both arms route the same network;
``muskingum_params`` / ``lateral_volumes`` repeat river_route_b200/synth.py (held equal by tests/test_refarm_cpu.py).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import subprocess
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
HOSTUTIL = os.path.join(HERE, '_build', 'librr_hostutil.so')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'river_route', 'routers', '_numba_kernels.py'))


def import_reference():
    """``import river_route`` from oracle/_ref with empty stand-ins for the I/O packages this image lacks (they are
    touched only by annotations / file I/O, never by the routing path)."""
    if 'river_route' in sys.modules:
        return sys.modules['river_route']
    os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(tempfile.gettempdir(), 'rr_refarm_numba_cache'))
    sys.dont_write_bytecode = True

    stubs = []

    def stub(name, **attrs):
        if name in sys.modules:
            return
        try:
            __import__(name)
            return
        except Exception:
            pass
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        stubs.append(name)

    stub('xarray', Dataset=type('Dataset', (), {}), DataArray=type('DataArray', (), {}))
    stub('netCDF4')
    stub('geopandas', GeoDataFrame=type('GeoDataFrame', (), {}))
    stub('shapely')
    stub('shapely.geometry', Point=object, MultiPoint=object, box=object)
    stub('shapely.ops', voronoi_diagram=object)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import river_route
    for name in stubs:            # needed at import time only; other code in this process must not find them
        sys.modules.pop(name, None)
    return river_route


# ---- synthetic inputs without the product library -------------------------------------------------------------
def _hostutil():
    if not os.path.exists(HOSTUTIL):
        subprocess.run(['make', '-s', '-C', HERE], check=True)
    lib = C.CDLL(HOSTUTIL)
    i32p, i64 = C.POINTER(C.c_int32), C.c_int64
    lib.rr_synth_forest.restype = C.c_int
    lib.rr_synth_forest.argtypes = [i64, i64, C.c_uint64, C.c_double, i64, C.c_double, i32p]
    lib.rr_label_basins.restype = C.c_int
    lib.rr_label_basins.argtypes = [i64, i32p, i32p, C.POINTER(i64), C.c_int32, i32p]
    return lib


def forest(n, n_basins=1, seed=0, depth_bias=0.5, main_stem=0, sigma=1.5):
    down = np.empty(n, dtype=np.int32)
    rc = _hostutil().rr_synth_forest(int(n), int(n_basins), int(seed), float(depth_bias), int(main_stem), float(sigma),
                                     down.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise RuntimeError('rr_synth_forest failed')
    return down


def basin_parts(down, n_parts):
    """LPT bin-packing of whole basins over n_parts workers (same routine as river_route_b200.label_basins)."""
    down = np.ascontiguousarray(down, dtype=np.int32)
    basin = np.empty(down.shape[0], dtype=np.int32)
    part = np.empty(down.shape[0], dtype=np.int32)
    nb = C.c_int64(0)
    p = C.POINTER(C.c_int32)
    rc = _hostutil().rr_label_basins(down.shape[0], down.ctypes.data_as(p), basin.ctypes.data_as(p), C.byref(nb),
                                     int(n_parts), part.ctypes.data_as(p))
    if rc:
        raise RuntimeError('rr_label_basins failed')
    return part


def muskingum_params(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(1800.0, 20000.0, n), rng.uniform(0.05, 0.4, n)


def lateral_volumes(T, n, seed=0):
    rng = np.random.default_rng(seed)
    out = np.empty((T, n), dtype=np.float64)
    step = max(1, (1 << 24) // max(n, 1))
    for t0 in range(0, T, step):
        t1 = min(T, t0 + step)
        blk = rng.gamma(0.3, 5.0e4, size=(t1 - t0, n))
        blk[rng.random((t1 - t0, n)) < 0.5] = 0.0
        out[t0:t1, :n] = blk
    return out


def local_network(down, idx):
    """Downstream index vector of the reaches ``idx`` (whole basins, original relative order) in local numbering."""
    new_of_old = np.full(down.shape[0], -1, dtype=np.int64)
    new_of_old[idx] = np.arange(idx.shape[0])
    d = down[idx]
    return np.where(d >= 0, new_of_old[np.where(d >= 0, d, 0)], -1)


# ---- the reference's router object, fed arrays instead of files ------------------------------------------------
def make_router(down_local, k, x, dt_runoff, dt_routing, rows):
    """A river_route.RapidMuskingum instance in the state its own ``route()`` reaches just before ``_router`` is
    called (Muskingum.py:199-227, TransformMuskingum.py:66-106), built with the reference's own functions; only the
    file reads (params parquet, qlateral netCDF) are replaced by arrays."""
    rr = import_reference()
    from river_route.tools import adjacency_matrix
    self = object.__new__(rr.RapidMuskingum)
    self.logger = logging.getLogger('river_route.refarm')
    self.logger.disabled = True
    n = down_local.shape[0]
    self.river_ids = np.arange(1, n + 1, dtype=np.int64)
    downstream_ids = np.where(down_local >= 0, down_local.astype(np.int64) + 1, -1)
    self.k, self.x = np.ascontiguousarray(k), np.ascontiguousarray(x)
    self.A = adjacency_matrix(self.river_ids, downstream_ids)                  # Muskingum.py:167
    self.dt_runoff, self.dt_routing = int(dt_runoff), int(dt_routing)
    self.num_runoff_steps = int(rows)
    self.num_routing_steps_per_runoff = int(dt_runoff / dt_routing)            # TransformMuskingum.py:99-101
    self._set_muskingum_coefficients(self.dt_routing)                         # Muskingum.py:172-193
    self.c4 = self.c1 + self.c2                                                # TransformMuskingum.py:104
    self.channel_state = np.zeros(n, dtype=np.float64)                         # Muskingum.py:119-121
    return self


def route_once(router, qlateral):
    """One file's worth of ``TransformMuskingum._execute_routing`` in sequential mode (:119-123)."""
    q_t, q_array = router._router(qlateral)                                    # RapidMuskingum.py:19-33
    router.channel_state = q_t
    return q_array


# ---- one core ---------------------------------------------------------------------------------------------------
def single_core(down, k, x, target_reaches, rows, dt=3600, ql=None):
    """rapid_route (numba, warm) on one core: the first whole basins of the network x ``rows`` steps."""
    outlets = np.flatnonzero(down < 0)
    m = int(outlets[np.searchsorted(outlets, min(target_reaches, down.shape[0]) - 1)]) + 1
    router = make_router(down[:m].astype(np.int64), k[:m], x[:m], dt, dt, rows)
    if ql is None:
        ql = lateral_volumes(rows, m, 99)
    warm_up()                                                                  # JIT compile / load the cache
    t = time.perf_counter()
    route_once(router, ql)
    sec = time.perf_counter() - t
    return {'value': m * rows / sec, 'reaches': m, 'rows': rows, 'seconds': sec}


# ---- all cores: independent watersheds in separate processes ----------------------------------------------------
_W = {}


def warm_up():
    """Import the reference and JIT-compile (or load from numba's cache) rapid_route for the runtime signature."""
    k, x = muskingum_params(8, 0)
    r = make_router(np.array([1, 2, 3, -1, 5, 6, 7, -1], dtype=np.int64), k, x, 3600, 3600, 2)
    route_once(r, np.zeros((2, 8)))


def _worker_init(rows, dt, counter):
    down, k, x, part = _W['shared']                       # inherited through fork, not pickled
    with counter.get_lock():
        me = counter.value
        counter.value += 1
    idx = np.flatnonzero(part == me)
    _W['id'] = me
    _W['n'] = int(idx.shape[0])
    if idx.shape[0] == 0:
        return
    _W['router'] = make_router(local_network(down, idx), k[idx], x[idx], dt, dt, rows)
    _W['ql'] = lateral_volumes(rows, idx.shape[0], 100 + me)
    _W['shared'] = None


def _worker_step(_):
    # every worker takes exactly one task per step: a worker holds its task until all of them have one
    _W['barrier'].wait()
    if _W['n'] == 0:
        return 0.0
    out = route_once(_W['router'], _W['ql'])
    return float(out[-1].sum())


class WatershedPool:
    """``cores`` processes, each owning the whole basins LPT-packed to it, its reference router object and its
    lateral inflows; one ``step()`` routes ``rows`` time steps of the whole network (state chained between steps),
    the way docs/references/parallelism.md:67-75 describes: watersheds are self-contained units, routed concurrently
    in separate processes."""

    def __init__(self, down, k, x, rows, cores, dt=3600):
        import multiprocessing as mp
        warm_up()                                          # compiled once in the parent, inherited by the workers
        ctx = mp.get_context('fork')
        self.cores = cores
        _W['shared'] = (down, k, x, basin_parts(down, cores))
        _W['barrier'] = ctx.Barrier(cores)
        counter = ctx.Value('i', 0)
        self.pool = ctx.Pool(cores, initializer=_worker_init, initargs=(rows, dt, counter))
        _W['shared'] = None
        self.step()                                        # returns once every worker has finished its initialiser

    def step(self):
        return sum(self.pool.map(_worker_step, range(self.cores), chunksize=1))

    def close(self):
        self.pool.close()
        self.pool.join()
