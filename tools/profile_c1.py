#!/usr/bin/env python
"""C1 with 12 routing substeps per row (50k reaches x 2920 rows): per-kernel-class device time for a few tile lengths and
stagings.  One JSON line per variant."""
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
n, T, dt_runoff = 50_000, 2920, 10800
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
down = synth.forest(n, 1, seed=0, depth_bias=0.9)
k, x = synth.muskingum_params(n, 0)
a = network_arrays(down, k, x, dt_runoff // K, dt_runoff)
d_lat = torch.from_numpy(synth.lateral_volumes(T, n, 0)).to(dev)
d_out = torch.empty((T, n), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
VARIANTS = [('auto', 0, 0), ('auto', 128, 0), ('auto', 256, 0), ('auto', 0, 6), ('auto', 0, 8), ('registers-tiled', 0, 0)]
if K > 1:
    VARIANTS += [('direct-nohw', 0, 0), ('direct-nohw', 128, 0), ('direct-nohw', 256, 0), ('direct-nohw', 0, 8)]
for staging, tile, stride in VARIANTS:
    plan = rr.Plan(down, staging=staging, time_tile=tile, tile_stride=stride)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    res = []
    for rep in range(3):
        d_q = torch.zeros(n, dtype=torch.float64, device=dev)
        timing_enable(True)
        timing_read(reset=True)
        plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), n, d_out.data_ptr(), n, T, K, stream)
        torch.cuda.synchronize()
        t = timing_read(reset=True)
        res.append({c: round(v['ms'], 3) for c, v in t.items()})
    timing_enable(False)
    print(json.dumps({'config': 'C1', 'substeps': K, 'staging': staging, 'time_tile': tile, 'tile_stride': stride, 'tile_rows': plan.tile_rows(T, K),
                      'ms_by_class_last_rep': res[-1], 'checksum': float(d_out.sum().item())}), flush=True)
    plan.close()
