#!/usr/bin/env python
"""Latency model of the wavefront on a small deep network (C1: 50k reaches, depth 189): route-kernel time against the
number of rows (16-row groups) -- slope = cost per group, intercept = cost of the levels.  One JSON line per variant."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth  # noqa: E402
from river_route_b200.plan import timing_enable, timing_read  # noqa: E402
from tests.helpers import network_arrays  # noqa: E402

dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'c1'
if which == 'c1':
    n, down = 50_000, synth.forest(50_000, 1, seed=0, depth_bias=0.9)
else:
    n, down = 500_000, synth.forest(500_000, 2, seed=1, depth_bias=0.5, main_stem=3000)
k, x = synth.muskingum_params(n, 0)
a = network_arrays(down, k, x, 3600, 3600)
Tmax = 2944
d_lat = torch.from_numpy(synth.lateral_volumes(64, n, 0)).to(dev).repeat(Tmax // 64, 1)
d_out = torch.empty((Tmax, n), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
plan = rr.Plan(down)
plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
info = plan.info
res = {'network': which, 'blocks': info['n_blocks'], 'levels': info['max_block_level'] + 1, 'spin_ns': os.environ.get('RR_PROG_SPIN_NS', '32'),
       'narrow': os.environ.get('RR_NARROW_BLOCKS', 'default')}
for T in (64, 256, 1024, 2944):
    best = 1e9
    for rep in range(3):
        d_q = torch.zeros(n, dtype=torch.float64, device=dev)
        timing_enable(True)
        timing_read(reset=True)
        plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), n, d_out.data_ptr(), n, T, 1, stream)
        torch.cuda.synchronize()
        best = min(best, timing_read(reset=True)['route']['ms'])
    res[f'T{T}_route_ms'] = round(best, 3)
timing_enable(False)
res['us_per_group'] = round((res['T2944_route_ms'] - res['T256_route_ms']) * 1e3 / ((2944 - 256) / 16), 2)
res['us_per_level_at_T64'] = round(res['T64_route_ms'] * 1e3 / res['levels'], 2)
print(json.dumps(res), flush=True)
