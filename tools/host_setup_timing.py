#!/usr/bin/env python
"""Host-side setup of the C4 configuration (7M reaches, 5000 basins): wall time of every step a router
does once per network, on this machine's CPU.  No GPU needed.  One JSON object (committed as profiles/r02_host_setup.json)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth  # noqa: E402
from river_route_b200.plan import downstream_index, label_basins  # noqa: E402
from river_route_b200.runoff import build_weight_csr  # noqa: E402
from river_route_b200.sharding import shard_by_basin  # noqa: E402

n, basins = int(sys.argv[1]) if len(sys.argv) > 1 else 7_000_000, 5000
res = {'reaches': n, 'basins': basins, 'cpu_threads_visible': os.cpu_count()}


def timed(name, fn):
    t = time.perf_counter()
    out = fn()
    res[name + '_s'] = round(time.perf_counter() - t, 3)
    return out


down = timed('synthetic_network (not a router step)', lambda: synth.forest(n, basins, seed=4, depth_bias=0.5))
rng = np.random.default_rng(0)
ids = rng.permutation(np.arange(10_000_000, 10_000_000 + n, dtype=np.int64))        # arbitrary, unsorted river ids
down_ids = np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1)
d2 = timed('downstream_index (ids -> indices, the three reference checks; tools.py:75-109)', lambda: downstream_index(ids, down_ids))
assert np.array_equal(d2, down)
timed('label_basins (union-find over the downstream links)', lambda: label_basins(down))
timed('shard_by_basin over 8 ranks (LPT by reach count)', lambda: shard_by_basin(down, 8, 0))
k, x = synth.muskingum_params(n, 4)
plan = timed('Plan (upstream CSR, levels, padded level-sorted order, block DAG, ticket metadata)', lambda: rr.Plan(down))
res['plan_info'] = {k_: v for k_, v in plan.info.items() if k_ in ('n_blocks', 'max_block_level', 'reach_depth', 'narrow_blocks', 'n_work', 'max_indegree')}
# weight table: ~1.7 cells per reach (12M rows), rivers in params order, cells on a 3600 x 1800 grid
rows = int(1.7 * n)
riv = np.sort(rng.integers(0, n, rows))
tab = dict(river_id=ids[riv], x_index=rng.integers(0, 3600, rows), y_index=rng.integers(0, 1800, rows), proportion=rng.random(rows),
           area_sqm=rng.random(rows) * 1e6)
out = timed(f'build_weight_csr ({rows} table rows; runoff.py:255-295)', lambda: build_weight_csr(**tab))
res['weight_csr'] = {'rivers': int(out[0].shape[0] - 1), 'entries': int(out[1].shape[0]), 'cells': int(out[3].shape[0])}
print(json.dumps(res, indent=1))
