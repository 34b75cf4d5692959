#!/usr/bin/env python
"""How much the kind of host memory matters for the router-level call (rr_route_host_ex): pinned vs pageable
numpy arrays, fresh vs touched output pages, and cudaHostRegister on the caller's arrays.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth  # noqa: E402
from tools.configs_report import bench_coefficients  # noqa: E402

n, T = 1_000_000, 256
down = synth.forest(n, 400, seed=2, depth_bias=0.5)
k, x = synth.muskingum_params(n, 2)
plan = rr.Plan(down)
plan.set_coefficients(*bench_coefficients(k, x))
lat_pin = rr.pinned_empty((T, n))
synth.lateral_volumes(32, n, 1, out=lat_pin[:32])
for r in range(32, T, 32):
    lat_pin[r:r + 32] = lat_pin[:32]
lat_page = np.array(lat_pin)                      # ordinary (pageable) numpy memory, pages touched
out_pin = rr.pinned_empty((T, n), dtype=np.float32)
q = np.zeros(n)
plan.route_host(rr.MODE_RAPID, q, lat_pin, out_pin, 1)
res = {}


def run(name, lat, make_out, reps=3):
    ts = []
    for _ in range(reps):
        q[:] = 0
        t = time.perf_counter()
        out = make_out()
        plan.route_host(rr.MODE_RAPID, q, lat, out, 1)
        ts.append(time.perf_counter() - t)
    res[name] = {'s': float(np.median(ts)), 'reach_steps_per_s': n * T / float(np.median(ts))}
    return out


ref = run('pinned_in_pinned_out', lat_pin, lambda: out_pin).copy()
touched = np.zeros((T, n), dtype=np.float32)
o = run('pageable_in_touched_out', lat_page, lambda: touched)
assert np.array_equal(o, ref)
run('pageable_in_fresh_out', lat_page, lambda: np.empty((T, n), dtype=np.float32))
run('pageable_in_pinned_alloc_out', lat_page, lambda: rr.pinned_empty((T, n), dtype=np.float32))
run('pinned_in_fresh_out', lat_pin, lambda: np.empty((T, n), dtype=np.float32))
rt = torch.cuda.cudart()
t = time.perf_counter()
rc = rt.cudaHostRegister(lat_page.ctypes.data, lat_page.nbytes, 0)
res['host_register_in'] = {'s': time.perf_counter() - t, 'rc': int(rc), 'GBps': lat_page.nbytes / 1e9 / (time.perf_counter() - t)}
run('registered_in_pinned_out', lat_page, lambda: out_pin)
t = time.perf_counter()
rt.cudaHostUnregister(lat_page.ctypes.data)
res['host_unregister_in'] = {'s': time.perf_counter() - t}
t = time.perf_counter()
tmp = rr.pinned_empty((T, n), dtype=np.float32)
res['pinned_alloc_1GB'] = {'s': time.perf_counter() - t}
t = time.perf_counter()
tmp2 = np.zeros((T, n), dtype=np.float32); tmp2[::1, ::1024] = 1
res['pageable_touch_1GB'] = {'s': time.perf_counter() - t}
t = time.perf_counter()
np.copyto(tmp, touched)
res['memcpy_1GB_one_thread'] = {'s': time.perf_counter() - t, 'GBps': tmp.nbytes / 1e9 / (time.perf_counter() - t)}
print(json.dumps({'probe': 'host memory kinds, 1M reaches x 256 steps, fp64 in / float32 out', **res}))
