"""Diagnostic: per-item phase cycles (RR_PROFILE build) for the C1 network with K substeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import river_route_b200 as rr
from river_route_b200 import synth
from tests.helpers import network_arrays
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n, T, dt_runoff = 50_000, 2920, 10800
down = synth.forest(n, 1, seed=0, depth_bias=0.9)
k, x = synth.muskingum_params(n, 0)
a = network_arrays(down, k, x, dt_runoff // K, dt_runoff)
plan = rr.Plan(down, **({'time_tile': int(sys.argv[2])} if len(sys.argv) > 2 else {}))
plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
dev = torch.device('cuda:0')
d_lat = torch.from_numpy(synth.lateral_volumes(T, n, 0)).to(dev)
d_out = torch.empty((T, n), dtype=torch.float64, device=dev)
for rep in range(2):
    d_q = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    if rep: plan.read_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), n, d_out.data_ptr(), n, T, K, torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    c = plan.read_profile(); items = max(c[6], 1)
    print(f'K={K} ms={e0.elapsed_time(e1):.1f} items={items} per-item cycles: ticket {c[0]/items:.0f} open {c[1]/items:.0f} [waits {c[4]/items:.0f}] body {c[2]/items:.0f} publish {c[3]/items:.0f}; general items {c[5]} avg body cycles {c[7]/max(c[5],1):.0f}', plan.info['max_block_level'])
