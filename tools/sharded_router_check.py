#!/usr/bin/env python
"""
N > 1 on hardware through the PRODUCT entry point: a basin-sharded RapidMuskingum run (one process per GPU, launched
with torchrun through `python -m river_route_b200.distributed`) must write the same discharge files and final state as a
single-GPU `RapidMuskingum(cfg).route()` of the same config.  One JSON line.

    python tools/sharded_router_check.py --gpus 2 [--reaches 400000 --rows 96]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=2)
    ap.add_argument('--reaches', type=int, default=400_000)
    ap.add_argument('--basins', type=int, default=300)
    ap.add_argument('--rows', type=int, default=96)
    ap.add_argument('--files', type=int, default=2)
    args = ap.parse_args()
    import river_route_b200 as rr
    from river_route_b200 import ncio, synth
    n, T = args.reaches, args.rows
    tmp = tempfile.mkdtemp(prefix='rr_sharded_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    down = synth.forest(n, args.basins, seed=4, depth_bias=0.5)
    k, x = synth.muskingum_params(n, 4)
    ids = np.arange(n, dtype=np.int64) + 1
    pd.DataFrame({'river_id': ids, 'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1),
                  'k': k, 'x': x}).to_parquet(os.path.join(tmp, 'params.parquet'))
    pd.DataFrame({'Q': np.random.default_rng(1).uniform(0, 30, n)}).to_parquet(os.path.join(tmp, 'q0.parquet'))
    files = []
    for f in range(args.files):
        path = os.path.join(tmp, f'ql_{f}.nc')
        ql = synth.lateral_volumes(T, n, 10 + f).astype(np.float32)
        with ncio.open_nc(path, 'w') as nc:
            nc.createDimension('time', T)
            nc.createDimension('river_id', n)
            tv = nc.createVariable('time', 'f8', ('time',))
            tv.units = 'seconds since 2022-01-01 00:00:00'
            tv[:] = (np.arange(T) + f * T) * 3600.0
            nc.createVariable('river_id', 'i4', ('river_id',))[:] = ids.astype(np.int32)
            nc.createVariable('qlateral', 'f4', ('time', 'river_id'))[:] = ql
        files.append(path)
    cfg = dict(params_file=os.path.join(tmp, 'params.parquet'), qlateral_files=files,
               channel_state_init_file=os.path.join(tmp, 'q0.parquet'), log=False)
    import yaml
    for label in ('single', 'sharded'):
        os.makedirs(os.path.join(tmp, label))
        with open(os.path.join(tmp, f'{label}.yaml'), 'w') as fh:
            yaml.safe_dump(dict(cfg, discharge_dir=os.path.join(tmp, label),
                                channel_state_final_file=os.path.join(tmp, label, 'final.parquet')), fh)
    t0 = time.perf_counter()
    rr.RapidMuskingum(os.path.join(tmp, 'single.yaml')).route()
    t_single = time.perf_counter() - t0
    t0 = time.perf_counter()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
           '--master-port', '29581', '-m', 'river_route_b200.distributed', 'RapidMuskingum', os.path.join(tmp, 'sharded.yaml')]
    run = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    t_sharded = time.perf_counter() - t0
    res = {'check': 'basin-sharded router run vs single-GPU run', 'gpus': args.gpus, 'reaches': n, 'rows_per_file': T,
           'files': args.files, 'torchrun_rc': run.returncode, 'single_s': t_single, 'sharded_s_incl_launch': t_sharded}
    if run.returncode != 0:
        res['stderr_tail'] = run.stderr[-1500:]
        print(json.dumps(res))
        return 1
    equal = True
    for f in range(args.files):
        with ncio.open_nc(os.path.join(tmp, 'single', f'discharge_ql_{f}.nc')) as a, \
                ncio.open_nc(os.path.join(tmp, 'sharded', f'discharge_ql_{f}.nc')) as b:
            for var in ('Q', 'river_id', 'time'):
                equal &= bool(np.array_equal(ncio.read_array(a.variables[var]), ncio.read_array(b.variables[var])))
    s1 = pd.read_parquet(os.path.join(tmp, 'single', 'final.parquet'))['Q'].values
    s2 = pd.read_parquet(os.path.join(tmp, 'sharded', 'final.parquet'))['Q'].values
    res.update(discharge_files_bitwise_equal=equal, final_state_bitwise_equal=bool(np.array_equal(s1, s2)))
    print(json.dumps(res))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return 0 if (equal and np.array_equal(s1, s2)) else 1


if __name__ == '__main__':
    sys.exit(main())
