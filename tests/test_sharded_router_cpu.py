"""
Basin-sharded router runs on the CPU tier: two gloo ranks (river_route_b200.distributed.Shard) each route the
drainage basins packed to them -- device entry points replaced by the oracle-backed stand-ins of tests/fakes.py -- and
rank 0 writes ONE discharge file and ONE state file that must equal, bit for bit, what a single unsharded run writes.
Covers qlateral files (RapidMuskingum, two files chained through the state) and grid files with a unit hydrograph
(UnitMuskingum: weight-table rows, UH kernel / state columns and the UH carry-over are sharded and gathered too).
"""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import scipy.sparse
import torch.multiprocessing as mp

import river_route_b200 as rr
from river_route_b200 import ncio, synth
from river_route_b200.distributed import Shard
from river_route_b200.runoff import QlateralDataset
from tests import fakes
from tests.test_routers_gpu import _grid_case

WORLD = 2


def _rapid_case(tmp, n=700, T=6):
    down = synth.forest(n, 9, seed=31, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 31)
    ids = np.arange(n, dtype=np.int64) * 3 + 11
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(os.path.join(tmp, 'p.parquet'))
    pd.DataFrame({'Q': np.random.default_rng(3).uniform(0, 20, n)}).to_parquet(os.path.join(tmp, 'q0.parquet'))
    files = []
    for f in range(2):
        t = (np.datetime64('2021-05-01') + (np.arange(T) + f * T) * np.timedelta64(1, 'h')).astype('datetime64[s]')
        QlateralDataset(synth.lateral_volumes(T, n, 40 + f), ids, t, 'm3').to_netcdf(os.path.join(tmp, f'ql_{f}.nc'))
        files.append(os.path.join(tmp, f'ql_{f}.nc'))
    return dict(params_file=os.path.join(tmp, 'p.parquet'), qlateral_files=files,
                channel_state_init_file=os.path.join(tmp, 'q0.parquet'), log=False)


def _run(kind, tmp, out, shard):
    os.makedirs(out, exist_ok=True)
    if kind == 'rapid':
        cfg = dict(_rapid_case.cfg, discharge_dir=out, channel_state_final_file=os.path.join(out, 'final.parquet'))
        rr.RapidMuskingum(_shard=shard, **cfg).route()
    else:
        cfg = dict(_run.unit_cfg, discharge_dir=out, channel_state_final_file=os.path.join(out, 'final.parquet'),
                   uh_state_final_file=os.path.join(out, 'uh1.parquet'))
        rr.UnitMuskingum(_shard=shard, **cfg).route()


def _worker(rank, port, tmp, kind, cfg):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    fakes.install(setattr)
    _rapid_case.cfg = cfg
    _run.unit_cfg = cfg
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    try:
        shard = Shard(rank, WORLD, gather_rows=4)          # several gather chunks
        _run(kind, tmp, os.path.join(tmp, 'sharded'), shard)
        shard.barrier()
    finally:
        dist.destroy_process_group()


def _spawn(tmp, kind, cfg):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(port, tmp, kind, cfg), nprocs=WORLD, join=True)


def _same_files(a, b, names):
    for nm in names:
        if nm.endswith('.parquet'):
            assert np.array_equal(pd.read_parquet(os.path.join(a, nm)).to_numpy(), pd.read_parquet(os.path.join(b, nm)).to_numpy()), nm
        else:
            with ncio.open_nc(os.path.join(a, nm)) as x, ncio.open_nc(os.path.join(b, nm)) as y:
                for var in ('Q', 'river_id', 'time'):
                    assert np.array_equal(ncio.read_array(x.variables[var]), ncio.read_array(y.variables[var])), (nm, var)


def test_rapid_two_ranks_write_what_one_process_writes(tmp_path, monkeypatch):
    tmp = str(tmp_path)
    cfg = _rapid_case(tmp)
    fakes.install(monkeypatch.setattr)
    _rapid_case.cfg = cfg
    _run('rapid', tmp, os.path.join(tmp, 'single'), None)
    _spawn(tmp, 'rapid', cfg)
    _same_files(os.path.join(tmp, 'single'), os.path.join(tmp, 'sharded'),
                ['discharge_ql_0.nc', 'discharge_ql_1.nc', 'final.parquet'])
    with ncio.open_nc(os.path.join(tmp, 'sharded', 'discharge_ql_1.nc')) as ds:
        assert ncio.read_array(ds.variables['Q']).shape == (6, 700)


def test_unit_from_grids_two_ranks(tmp_path, monkeypatch):
    tmp = str(tmp_path)
    c = _grid_case(tmp_path, n=500, T=10, ny=8, nx=11)
    rng = np.random.default_rng(9)
    ker = rng.uniform(0, 1, (5, c['n'])) * (rng.random((5, c['n'])) < 0.7)
    scipy.sparse.save_npz(os.path.join(tmp, 'uh.npz'), scipy.sparse.csr_matrix(ker))
    s0 = rng.uniform(0, 1e-3, ker.shape)
    s0[-1] = 0
    pd.DataFrame(s0.T).to_parquet(os.path.join(tmp, 'uh0.parquet'))
    cfg = dict(params_file=c['params'], grid_runoff_files=[g[0] for g in c['grids']], grid_weights_file=os.path.join(tmp, 'weights.nc'),
               channel_state_init_file=c['state'], uh_kernel_file=os.path.join(tmp, 'uh.npz'),
               uh_state_init_file=os.path.join(tmp, 'uh0.parquet'), var_x='lon', var_y='lat', log=False)
    fakes.install(monkeypatch.setattr)
    _run.unit_cfg = cfg
    _run('unit', tmp, os.path.join(tmp, 'single'), None)
    _spawn(tmp, 'unit', cfg)
    names = [f'discharge_{os.path.basename(g[0])}' for g in c['grids']] + ['final.parquet', 'uh1.parquet']
    _same_files(os.path.join(tmp, 'single'), os.path.join(tmp, 'sharded'), names)


def test_shard_rejects_output_subsets(tmp_path, monkeypatch):
    cfg = _rapid_case(str(tmp_path))
    fakes.install(monkeypatch.setattr)
    r = rr.RapidMuskingum(_shard=Shard(0, 1), discharge_dir=str(tmp_path), **cfg).set_output_rivers([11])
    with pytest.raises(ValueError, match='basin-sharded'):
        r.route()
