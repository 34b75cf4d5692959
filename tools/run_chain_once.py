#!/usr/bin/env python
"""C1 / C2 (small deep networks) routed a few times on the device: the command line ncu wraps for the narrow-level path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth  # noqa: E402
from tests.helpers import network_arrays  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'c1'
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2944
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda:0')
if which == 'c1':
    n, down = 50_000, synth.forest(50_000, 1, seed=0, depth_bias=0.9)
else:
    n, down = 500_000, synth.forest(500_000, 2, seed=1, depth_bias=0.5, main_stem=3000)
k, x = synth.muskingum_params(n, 0)
a = network_arrays(down, k, x, 3600, 3600)
d_lat = torch.from_numpy(synth.lateral_volumes(64, n, 0)).to(dev).repeat((T + 63) // 64, 1)[:T].contiguous()
d_out = torch.empty((T, n), dtype=torch.float64, device=dev)
plan = rr.Plan(down)
plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
stream = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    d_q = torch.zeros(n, dtype=torch.float64, device=dev)
    plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), n, d_out.data_ptr(), n, T, 1, stream)
    torch.cuda.synchronize()
print('checksum', float(d_out.sum()))
