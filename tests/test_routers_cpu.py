"""Host logic of the router classes that needs no GPU: the config surface and the errors of the reference."""
import json
import os

import numpy as np
import pandas as pd
import pytest

import river_route_b200 as rr
from river_route_b200 import synth


def _params(tmp_path, n=50, shuffle_ids=True):
    down = synth.forest(n, 2, seed=3, depth_bias=0.6)
    ids = (np.random.default_rng(0).permutation(n) + 1) * 3 if shuffle_ids else np.arange(1, n + 1)
    k, x = synth.muskingum_params(n, 1)
    path = str(tmp_path / 'params.parquet')
    pd.DataFrame({'river_id': ids.astype(np.int64),
                  'downstream_river_id': np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1).astype(np.int64),
                  'k': k, 'x': x}).to_parquet(path)
    return path, down, ids


def test_config_surface_matches_reference_keys(tmp_path):
    params, _, _ = _params(tmp_path)
    keys = {'params_file', 'discharge_dir', 'discharge_files', 'channel_state_init_file', 'channel_state_final_file',
            'dt_routing', 'dt_total', 'dt_discharge', 'dt_runoff', 'start_datetime', 'qlateral_files',
            'grid_runoff_files', 'grid_weights_file', 'grid_accumulation_type', 'runoff_processing_mode',
            'uh_kernel_file', 'uh_state_init_file', 'uh_state_final_file', 'log', 'progress_bar', 'log_level',
            'log_stream', 'log_format', 'var_river_id', 'var_discharge', 'var_grid_runoff', 'var_x', 'var_y', 'var_t'}
    cfg = rr.Configs(params_file=params, discharge_dir=str(tmp_path))      # river_route/routers/Config.py:31-65
    assert keys == {f.name for f in __import__('dataclasses').fields(cfg)}
    assert cfg.discharge_files == [os.path.join(str(tmp_path), 'discharge.nc')]          # Config.py:161-163
    assert os.path.isabs(cfg.params_file)
    cfg = rr.Configs(params_file=params, discharge_dir=str(tmp_path), qlateral_files=params, log=False)
    assert cfg.qlateral_files == [params] and cfg.progress_bar is False                  # Config.py:97, :112-118
    assert cfg.discharge_files == [os.path.join(str(tmp_path), 'discharge_params.parquet')]   # Config.py:155-159


def test_config_errors(tmp_path):
    params, _, _ = _params(tmp_path)
    with pytest.raises(ValueError, match='discharge_dir'):
        rr.Configs(params_file=params)
    with pytest.raises(ValueError, match='not both'):
        rr.Configs(params_file=params, discharge_dir=str(tmp_path), discharge_files=[str(tmp_path / 'a.nc')])
    with pytest.raises(FileNotFoundError, match='params_file not found'):
        rr.Configs(params_file=str(tmp_path / 'missing.parquet'), discharge_dir=str(tmp_path))
    with pytest.raises(FileNotFoundError, match='qlateral_files'):
        rr.Configs(params_file=params, discharge_dir=str(tmp_path), qlateral_files=[str(tmp_path / 'nope.nc')])
    with pytest.raises(NotADirectoryError):
        rr.Configs(params_file=params, discharge_files=[str(tmp_path / 'no_dir' / 'q.nc')])
    with pytest.raises(ValueError, match='runoff_processing_mode must be one of'):
        rr.Configs(params_file=params, discharge_dir=str(tmp_path), runoff_processing_mode='parallel')
    with pytest.raises(ValueError, match='Missing required config: params_file'):
        rr.Configs(discharge_dir=str(tmp_path))
    with pytest.raises(TypeError):
        rr.Configs(params_file=params, discharge_dir=str(tmp_path), not_a_key=1)
    with pytest.raises(RuntimeError, match='Unrecognized simulation config file type'):   # Muskingum.py:78
        rr.Muskingum(str(tmp_path / 'config.toml'))


def test_config_files_and_kwarg_override(tmp_path):
    params, _, _ = _params(tmp_path)
    import yaml
    yml, jsn = tmp_path / 'c.yaml', tmp_path / 'c.json'
    base = dict(params_file=params, discharge_dir=str(tmp_path), dt_routing=900, dt_total=3600, log=False)
    yml.write_text(yaml.safe_dump(base))
    jsn.write_text(json.dumps(base))
    for f in (yml, jsn):
        r = rr.Muskingum(str(f), dt_routing=1800)          # kwargs win over the file (Muskingum.py:69-81)
        assert r.cfg.dt_routing == 1800 and r.cfg.dt_total == 3600


def test_router_validation_errors_fire_before_any_gpu_work(tmp_path):
    params, down, _ = _params(tmp_path)
    state = str(tmp_path / 'state.parquet')
    pd.DataFrame({'Q': np.ones(down.shape[0])}).to_parquet(state)
    common = dict(params_file=params, discharge_dir=str(tmp_path), log=False)
    with pytest.raises(ValueError, match='channel_state_init_file is required for Muskingum'):
        rr.Muskingum(**common, dt_routing=900, dt_total=3600).route()
    with pytest.raises(ValueError, match='dt_total must be an integer multiple of dt_discharge'):
        rr.Muskingum(**common, channel_state_init_file=state, dt_routing=900, dt_discharge=1800, dt_total=4500).route()
    with pytest.raises(ValueError, match='dt_discharge must be an integer multiple of dt_routing'):
        rr.Muskingum(**common, channel_state_init_file=state, dt_routing=700, dt_discharge=1800, dt_total=3600).route()
    with pytest.raises(ValueError, match='Provide qlateral_files or grid_runoff_files'):
        rr.RapidMuskingum(**common).route()
    with pytest.raises(ValueError, match='uh_kernel_file is required for UnitMuskingum'):
        rr.UnitMuskingum(**common, qlateral_files=[params]).route()

    class Fake(rr.RapidMuskingum):
        def _qlateral_generator(self):
            dates = np.datetime64('2020-01-01') + np.arange(4) * np.timedelta64(3600, 's')
            yield dates.astype('datetime64[s]'), np.zeros((4, self.n)), 'in', self.cfg.discharge_files[0]

    with pytest.raises(ValueError, match='dt_runoff must be an integer multiple of dt_routing'):
        Fake(**common, qlateral_files=[params], dt_routing=1000).route()
    with pytest.raises(ValueError, match='dt_runoff must be >= dt_routing'):
        Fake(**common, qlateral_files=[params], dt_routing=7200).route()


def test_topology_errors_through_the_router(tmp_path):
    n = 6
    bad = str(tmp_path / 'bad.parquet')
    pd.DataFrame({'river_id': [1, 2, 3, 4, 5, 6], 'downstream_river_id': [2, 3, 1, 5, 6, -1],
                  'k': np.full(n, 3000.0), 'x': np.full(n, 0.2)}).to_parquet(bad)
    state = str(tmp_path / 's.parquet')
    pd.DataFrame({'Q': np.ones(n)}).to_parquet(state)
    with pytest.raises(ValueError, match='topologically sorted'):
        rr.Muskingum(params_file=bad, discharge_dir=str(tmp_path), channel_state_init_file=state, dt_routing=900,
                     dt_total=3600, log=False).route()
    dup = str(tmp_path / 'dup.parquet')
    pd.DataFrame({'river_id': [1, 1, 3], 'downstream_river_id': [3, 3, -1], 'k': np.full(3, 3000.0),
                  'x': np.full(3, 0.2)}).to_parquet(dup)
    with pytest.raises(ValueError, match='duplicate river IDs'):
        rr.Muskingum(params_file=dup, discharge_dir=str(tmp_path), channel_state_init_file=state, dt_routing=900,
                     dt_total=3600, log=False).route()
