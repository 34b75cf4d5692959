#!/usr/bin/env python
"""
Times and parity-checks the BASELINE.json configurations C1-C3 at full size on one GPU (C4 is bench.py):
  C1 RapidMuskingum 50k reaches, 1 year 3-hourly (dt_routing = dt_runoff and the 12-substep variant)
  C2 Muskingum channel-only, 500k reaches, main stem 3000, 15 days at 900 s (1440 steps), Q0 = 10
  C3 UnitMuskingum 1M reaches: grid weights SpMM + unit-hydrograph convolution + routing, T = 744
Each stage is compared with the CPU oracle (full size where that takes seconds, a basin subset otherwise)
with the parity measure of SURVEY.md 8d.  Writes one JSON object per line.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import river_route_b200 as rr  # noqa: E402
from river_route_b200 import synth, _lib  # noqa: E402
from river_route_b200.plan import timing_enable, timing_read  # noqa: E402
from river_route_b200._lib import lib, check  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.helpers import network_arrays, parity_error  # noqa: E402
import ctypes as C  # noqa: E402

dev = torch.device('cuda:0')
PEAK = 6544.3


def emit(**kw):
    print(json.dumps(kw), flush=True)


def dev_route(plan, mode, q0, lat, T, K, n, reps=3):
    """Device-resident timing of one call (median of reps), returns (ms, out, q)."""
    d_lat = torch.from_numpy(lat).to(dev) if lat is not None else None
    d_out = torch.empty((T, n), dtype=torch.float64, device=dev)
    times = []
    for _ in range(reps):
        d_q = torch.from_numpy(q0).to(dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        plan.route_dev(mode, d_q.data_ptr(), d_lat.data_ptr() if d_lat is not None else 0, n, d_out.data_ptr(), n, T, K,
                       torch.cuda.current_stream().cuda_stream)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), d_out.cpu().numpy(), d_q.cpu().numpy()


def c1():
    n, T, dt_runoff = 50_000, 2920, 10800
    down = synth.forest(n, 1, seed=0, depth_bias=0.9)
    k, x = synth.muskingum_params(n, 0)
    ql = synth.lateral_volumes(T, n, 0)
    q0 = np.zeros(n)
    for dt_routing in (10800, 900):
        K = dt_runoff // dt_routing
        a = network_arrays(down, k, x, dt_routing, dt_runoff)
        plan = rr.Plan(down)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
        ms, out, q = dev_route(plan, rr.MODE_RAPID, q0, ql, T, K, n)
        t = time.perf_counter()
        q_ref, ref = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, K)
        cpu_s = time.perf_counter() - t
        h_out, h_q = np.empty((T, n)), q0.copy()
        t = time.perf_counter()
        plan.route_host(rr.MODE_RAPID, h_q, ql, h_out, K)
        e2e_s = time.perf_counter() - t
        emit(config='C1', router='RapidMuskingum', reaches=n, steps=T, substeps=K, depth=synth.depth(down),
             plan=plan.info, gpu_ms=ms, reach_substeps_per_s=n * T * K / (ms * 1e-3), e2e_pageable_s=e2e_s,
             cpu_oracle_s=cpu_s, cpu_reach_substeps_per_s=n * T * K / cpu_s, parity=parity_error(out, ref),
             parity_state=parity_error(q, q_ref), host_equals_dev=bool(np.array_equal(h_out, out)),
             clamp_pattern_equal=bool(np.array_equal(out == 0, ref == 0)))
        plan.close()


def c2():
    n, T = 500_000, 1440
    down = synth.forest(n, 2, seed=1, depth_bias=0.5, main_stem=3000)
    k, x = synth.muskingum_params(n, 1)
    a = network_arrays(down, k, x, 900)
    q0 = np.full(n, 10.0)
    for ren in ('auto', 'never'):
        plan = rr.Plan(down, renumber=ren)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
        ms, out, q = dev_route(plan, rr.MODE_MUSKINGUM, q0, None, T, 1, n)
        rec = dict(config='C2', router='Muskingum', renumber=ren, reaches=n, steps=T, depth=synth.depth(down),
                   plan=plan.info, gpu_ms=ms, reach_steps_per_s=n * T / (ms * 1e-3),
                   alg_GBps=n * T * 40 / (ms * 1e-3) / 1e9)
        if ren == 'auto':
            t = time.perf_counter()
            q_ref, ref = q0.copy(), np.zeros((T, n))
            oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_ref, ref, T, 1)
            rec.update(cpu_oracle_s=time.perf_counter() - t, parity=parity_error(out, ref),
                       parity_state=parity_error(q, q_ref), all_nonnegative=bool((out >= 0).all()))
        emit(**rec)
        plan.close()


def triangular_uh(k, area, tr):
    """SCS triangular unit hydrograph columns (test input only; the reference builds these offline)."""
    tp = 0.6 * 5.0 * k + tr / 2.0
    tb = 2.67 * tp
    nst = int(np.ceil(tb / tr).max())
    edges = np.minimum(np.arange(nst + 1)[:, None] * tr, tb[None, :])

    def cum(t):
        up = np.minimum(t, tp) ** 2 / (2 * tp)
        dn = np.where(t > tp, (t - tp) * 1.0 - (t - tp) ** 2 / (2 * (tb - tp)), 0.0)
        return (up + dn)
    u = cum(edges)
    ker = np.diff(u, axis=0) / tr
    ker *= area[None, :] / (ker.sum(axis=0) * tr)
    return ker


def c3():
    n, T, dt = 1_000_000, 744, 3600
    down = synth.forest(n, 400, seed=2, depth_bias=0.5)
    k, x = synth.muskingum_params(n, 2)
    rng = np.random.default_rng(2)
    a = network_arrays(down, k, x, dt, dt)
    # ---- grid weights: ERA5 0.25 degree grid, 4-8 cells per river ----
    ny, nx = 721, 1440
    ncell = rng.integers(4, 9, n)
    river_idx = np.repeat(np.arange(n), ncell)
    base = rng.integers(0, ny * nx - 3000, n)
    cells = np.repeat(base, ncell) + rng.integers(0, 3000, river_idx.shape[0])
    prop = rng.random(river_idx.shape[0]) + 0.05
    prop /= np.repeat(np.add.reduceat(prop, np.concatenate([[0], np.cumsum(ncell)[:-1]])), ncell)
    uniq, point_idx = np.unique(cells, return_inverse=True)
    indptr, indices, data = oracle.weights_csr(river_idx, point_idx, prop, n, uniq.shape[0])
    area = rng.uniform(1e5, 5e8, n)
    grid = rng.gamma(0.3, 2e-3, (T, uniq.shape[0])).astype(np.float32)
    grid[rng.random(grid.shape) < 0.6] = 0.0
    d_ptr, d_idx = torch.from_numpy(indptr).to(dev), torch.from_numpy(indices).to(dev)
    d_w, d_x = torch.from_numpy(data).to(dev), torch.from_numpy(grid).to(dev)
    d_y = torch.empty((T, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=3):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    vp = C.c_void_p
    ms_w = timed(lambda: check(lib.rr_weights_transform_dev(n, uniq.shape[0], T, vp(d_ptr.data_ptr()), vp(d_idx.data_ptr()), vp(d_w.data_ptr()),
                                                            vp(d_x.data_ptr()), 1, uniq.shape[0], vp(d_y.data_ptr()), n, 0, 0,
                                                            None, vp(stream))))
    depths = d_y.cpu().numpy()
    sub = slice(0, 20000)
    ref_w = oracle.weights_transform(indptr[:20001], indices[:indptr[20000]], data[:indptr[20000]], grid)
    bytes_w = 8 * n * T + indices.shape[0] * 12 + uniq.shape[0] * T * 4
    emit(config='C3', stage='weights_transform', rivers=n, nnz=int(indices.shape[0]), cells=int(uniq.shape[0]), steps=T,
         gpu_ms=ms_w, river_steps_per_s=n * T / (ms_w * 1e-3), alg_GBps=bytes_w / (ms_w * 1e-3) / 1e9,
         frac_of_hbm=bytes_w / (ms_w * 1e-3) / 1e9 / PEAK, parity_subset=parity_error(depths[:, sub], ref_w))
    # ---- unit hydrograph ----
    ker = triangular_uh(k, area, float(dt))
    n_ks = ker.shape[0]
    d_k = torch.from_numpy(np.ascontiguousarray(ker)).to(dev)
    d_s = torch.zeros((n_ks, n), dtype=torch.float64, device=dev)
    d_c = torch.empty((T, n), dtype=torch.float64, device=dev)

    def run_uh():
        d_s.zero_()
        check(lib.rr_uh_convolve_dev(n, n_ks, T, vp(d_y.data_ptr()), n, vp(d_k.data_ptr()), n, vp(d_s.data_ptr()), n,
                                     vp(d_c.data_ptr()), n, vp(stream)))
    ms_u = timed(run_uh)
    conv = d_c.cpu().numpy()
    st_ref = np.zeros((n_ks, 20000))
    ref_u = oracle.uh_convolve(depths[:, sub], ker[:, sub], st_ref)
    col = np.max(np.abs(ref_u), axis=0)
    emit(config='C3', stage='uh_convolve', basins=n, n_ks=int(n_ks), steps=T, gpu_ms=ms_u,
         basin_steps_per_s=n * T / (ms_u * 1e-3), alg_GBps=16 * n * T / (ms_u * 1e-3) / 1e9,
         frac_of_hbm=16 * n * T / (ms_u * 1e-3) / 1e9 / PEAK, fp64_tflops=2 * n_ks * n * T / (ms_u * 1e-3) / 1e12,
         parity_subset=parity_error(conv[:, sub], ref_u, col),
         parity_state_subset=parity_error(d_s[:, sub].cpu().numpy(), st_ref, col))
    # ---- unit route ----
    q0 = rng.uniform(0, 20, n)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
    ms_r, out, q = dev_route(plan, rr.MODE_UNIT, q0, conv, T, 1, n)
    sp = oracle.unit_split(down.astype(np.int64))
    inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
    c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
    q_ch = q0[inner].copy(); q_full = q_ch.copy(); ref = np.zeros((T, n))
    t = time.perf_counter()
    oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner, q_ch,
                      q_full, conv, ref, 1)
    cpu_s = time.perf_counter() - t
    emit(config='C3', stage='unit_route', reaches=n, inner=int(inner.shape[0]), headwater=int(hw.shape[0]), steps=T,
         plan=plan.info, gpu_ms=ms_r, reach_steps_per_s=n * T / (ms_r * 1e-3), cpu_oracle_s=cpu_s,
         cpu_reach_steps_per_s=n * T / cpu_s, parity=parity_error(out, ref))


if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c2', 'c3']
    for w in which:
        {'c1': c1, 'c2': c2, 'c3': c3}[w]()
