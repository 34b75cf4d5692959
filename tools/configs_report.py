#!/usr/bin/env python
"""
Times and parity-checks the BASELINE.json configurations at full size on one GPU (the C4 bench line is bench.py):
  C1 RapidMuskingum 50k reaches, 1 year 3-hourly (dt_routing = dt_runoff and the 12-substep variant)
  C2 Muskingum channel-only, 500k reaches, main stem 3000, 15 days at 900 s (1440 steps), Q0 = 10
  C3 UnitMuskingum 1M reaches: grid weights SpMM + unit-hydrograph convolution + routing, T = 744, stage by stage
     and as the single host call the router makes (rr_runoff_route_host: float32 grid in, float32 discharge out)
  C4 RapidMuskingum 7M reaches / 5000 basins: parity of a 240-step resident chunk against the oracle on whole
     basins (the oracle at full size would take minutes), then a full year (8760 hourly steps) streamed from
     pinned host memory through the router-level call, state chained from chunk to chunk
  C5 ensemble: 51 members x 360 hourly steps on the 7M-reach network, every member from the same initial state
     (TransformMuskingum.py:121-126), final state = member mean; parity of members on whole basins
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
from river_route_b200 import synth  # noqa: E402
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


def c3_pipeline():
    """C3 the way UnitMuskingum.route() runs it on grid files: one rr_runoff_route_host call per file."""
    from river_route_b200.transforms import Transform
    n, T, dt = 1_000_000, 744, 3600
    down = synth.forest(n, 400, seed=2, depth_bias=0.5)
    k, x = synth.muskingum_params(n, 2)
    rng = np.random.default_rng(2)
    a = network_arrays(down, k, x, dt, dt)
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
    ker = triangular_uh(k, area, float(dt))
    grid = rr.pinned_empty((T, uniq.shape[0]), dtype=np.float32)
    grid[:] = rng.gamma(0.3, 2e-3, grid.shape).astype(np.float32)
    grid[rng.random(grid.shape) < 0.6] = 0.0
    q0 = rng.uniform(0, 20, n)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
    tf = Transform(indptr, indices, data, uniq.shape[0], area=area).set_unit_hydrograph(ker)
    out = rr.pinned_empty((T, n), dtype=np.float32)
    q = q0.copy()
    plan.runoff_route_host(tf, rr.MODE_UNIT, q, grid, out, 1)                 # warm-up: allocations, plan upload
    times = []
    for _ in range(3):
        tf.set_unit_hydrograph(ker)
        q = q0.copy()
        timing_enable(True); timing_read(reset=True)
        t = time.perf_counter()
        plan.runoff_route_host(tf, rr.MODE_UNIT, q, grid, out, 1)
        times.append(time.perf_counter() - t)
        kt = timing_read(reset=True); timing_enable(False)
    # parity on whole basins (basins are contiguous index ranges in the generator's order)
    m = first_basins(down, 60_000)
    sub = slice(0, m)
    dep = oracle.weights_transform(indptr[:m + 1], indices[:indptr[m]], data[:indptr[m]], np.asarray(grid))
    st = np.zeros((ker.shape[0], m))
    conv = oracle.uh_convolve(dep, ker[:, sub], st)
    sd = down[sub]
    sa = network_arrays(sd, k[sub], x[sub], dt, dt)
    sp = oracle.unit_split(sd.astype(np.int64))
    inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
    c1i, c2i, c3i = sa['c1'][inner], sa['c2'][inner], sa['c3'][inner]
    q_ch = q0[sub][inner].copy(); q_full = q_ch.copy(); ref = np.zeros((T, m))
    oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner, q_ch,
                      q_full, conv, ref, 1)
    ref32 = ref.astype(np.float32)
    got = out[:, sub]
    rel = np.abs(got - ref32) / (np.abs(ref32) + np.abs(ref32).max(axis=0, keepdims=True) + 1e-30)
    s = float(np.median(times))
    emit(config='C3', stage='grid_to_discharge_one_call', api='rr_runoff_route_host (UNIT)', reaches=n, steps=T,
         cells=int(uniq.shape[0]), n_ks=int(ker.shape[0]), wall_s=s, reach_steps_per_s=n * T / s,
         h2d_bytes=int(grid.nbytes), d2h_bytes=int(out.nbytes), device_ms_by_class={k_: v['ms'] for k_, v in kt.items()},
         parity_reaches=m, float32_max_rel_diff=float(rel.max()), float32_ulp_ok=bool(rel.max() < 2e-7))
    tf.close(); plan.close()


def c4_network():
    n = 7_000_000
    down = synth.forest(n, 5000, seed=4, depth_bias=0.5)
    k, x = synth.muskingum_params(n, 4)
    return n, down, k, x


def first_basins(down, target):
    """Number of reaches in the first whole basins holding at least `target` reaches (basins are contiguous ranges)."""
    outlets = np.flatnonzero(down < 0)
    return int(outlets[min(np.searchsorted(outlets, target), outlets.shape[0] - 1)]) + 1


def subset_parity(down, k, x, m, lat_rows, out_rows, q0, q_after, dt=3600):
    """RapidMuskingum parity of the first m reaches (whole basins) against the oracle; arrays are already cut to m."""
    a = network_arrays(down[:m], k[:m], x[:m], dt, dt)
    q_ref, ref = np.ascontiguousarray(q0, dtype=np.float64).copy(), np.zeros((lat_rows.shape[0], m))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref,
                       np.ascontiguousarray(lat_rows), ref, 1)
    return parity_error(out_rows, ref), parity_error(q_after, q_ref)


def c4():
    n, down, k, x = c4_network()
    a1 = bench_coefficients(k, x)
    plan = rr.Plan(down)
    plan.set_coefficients(*a1)
    rows = 240
    lat = rr.pinned_empty((rows, n))
    synth.lateral_volumes(48, n, 5, out=lat[:48])                              # 48 distinct rows, repeated with a trend
    for r in range(48, rows, 48):
        np.multiply(lat[:48], 1.0 + 0.1 * (r // 48), out=lat[r:r + 48])
    q0 = np.zeros(n)
    ms, out, q = dev_route(plan, rr.MODE_RAPID, q0, lat, rows, 1, n, reps=3)
    m = first_basins(down, 300_000)
    p_out, p_q = subset_parity(down, k, x, m, lat[:, :m], out[:, :m], q0[:m], q[:m])
    emit(config='C4', stage='resident_chunk', reaches=n, steps=rows, gpu_ms=ms, reach_steps_per_s=n * rows / (ms * 1e-3),
         parity_reaches=m, parity=p_out, parity_state=p_q, checksum_out=float(out.sum()), checksum_state=float(q.sum()))
    del out
    # ---- a full year, streamed: 8760 hourly steps = 91.25 chunks of 96 rows cycled from one pinned buffer ----
    er, year = 96, 8760
    out32 = rr.pinned_empty((er, n), dtype=np.float32)
    qy = np.zeros(n)
    plan.route_host(rr.MODE_RAPID, qy, lat[:er], out32, 1)                     # warm-up
    qy[:] = 0.0
    t = time.perf_counter()
    done = 0
    while done < year:
        r = min(er, year - done)
        plan.route_host(rr.MODE_RAPID, qy, lat[:r], out32[:r], 1)
        done += r
    s = time.perf_counter() - t
    emit(config='C4', stage='one_year_streamed', api='Plan.route_host -> rr_route_host_ex (float32 out)', reaches=n,
         steps=year, wall_s=s, reach_steps_per_s=n * year / s, h2d_GB=n * year * 8 / 1e9, d2h_GB=n * year * 4 / 1e9,
         state_finite=bool(np.isfinite(qy).all()), state_sum=float(qy.sum()))
    # ---- the same year from gridded runoff (ERA5 0.25 degree, float32) to the hydrographs of the 5000 outlets:
    #      weight table -> route on the device, outlet columns copied back (rr_plan_set_output_subset) ----
    from river_route_b200.transforms import Transform
    n_cells = 721 * 1440
    rng = np.random.default_rng(77)
    per = rng.integers(4, 9, n)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(per, out=indptr[1:])
    first = rng.integers(0, n_cells - 8, n)
    indices = (np.repeat(first, per) + (np.arange(indptr[-1]) - np.repeat(indptr[:-1], per))).astype(np.int32)
    w = rng.random(indptr[-1])
    w /= np.repeat(np.add.reduceat(w, indptr[:-1]), per)
    tf = Transform(indptr, indices, w, n_cells, area=rng.uniform(1e5, 5e8, n))
    gr = 192
    grid = rr.pinned_empty((gr, n_cells), dtype=np.float32)
    grid[:] = (rng.gamma(0.3, 2e-3, (gr, n_cells)) * (rng.random((gr, n_cells)) < 0.4)).astype(np.float32)
    outlets = np.flatnonzero(down < 0).astype(np.int32)
    plan.set_output_subset(outlets)
    hyd = rr.pinned_empty((gr, outlets.shape[0]), dtype=np.float32)
    qy[:] = 0.0
    plan.runoff_route_host(tf, rr.MODE_RAPID, qy, grid, hyd, 1, as_volumes=True)       # warm-up
    qy[:] = 0.0
    t = time.perf_counter()
    done = 0
    while done < year:
        r = min(gr, year - done)
        plan.runoff_route_host(tf, rr.MODE_RAPID, qy, grid[:r], hyd[:r], 1, as_volumes=True)
        done += r
    s = time.perf_counter() - t
    emit(config='C4', stage='one_year_grid_to_outlet_hydrographs',
         api='Plan.set_output_subset + Plan.runoff_route_host (float32 grid in, float32 outlet discharge out)', reaches=n,
         outlets=int(outlets.shape[0]), steps=year, wall_s=s, reach_steps_per_s=n * year / s,
         h2d_GB=n_cells * year * 4 / 1e9, d2h_GB=outlets.shape[0] * year * 4 / 1e9, state_finite=bool(np.isfinite(qy).all()))
    plan.set_output_subset(None)
    tf.close()
    plan.close()


def bench_coefficients(k, x, dt=3600):
    dt_div_k = dt / k
    den = dt_div_k + (2 * (1 - x))
    _2x = 2 * x
    c1 = (dt_div_k - _2x) / den
    c2 = (dt_div_k + _2x) / den
    c3 = ((2 * (1 - x)) - dt_div_k) / den
    return c1, c2, c3, (c1 + c2) / dt


def c5():
    """Ensemble forecast, 51 members x 360 hourly steps x 7M reaches.  51 x 360 x 7M x 8 B = 1 TB of lateral inflows do
    not fit one GPU, so the members stream from (pinned) host memory whatever is done on the device; two figures:
    (1) device resident: 5 members x 120 rows routed by ONE batched call (rr_route_ensemble_dev: stage_in / stage_out
        per member, one wavefront launch over tickets x members), per-member parity on whole basins, against the same
        5 members routed one call each;
    (2) host arrays: rr_route_ensemble_host, 6 members x 96 rows of pinned float32 inflows in, float32 discharge out,
        members of a chunk in one launch, mean state on the device (TransformMuskingum.py:121-126, :145-146).
    Member inputs = one base series x lognormal(0, 0.3) per member."""
    n, down, k, x = c4_network()
    plan = rr.Plan(down)
    plan.set_coefficients(*bench_coefficients(k, x))
    stream = torch.cuda.current_stream().cuda_stream
    q_init = np.random.default_rng(6).uniform(0, 30, n)
    d_q0 = torch.from_numpy(q_init).to(dev)
    scale = np.random.default_rng(7).lognormal(0, 0.3, 51)
    base = torch.from_numpy(synth.lateral_volumes(24, n, 6)).to(dev)
    # ---- (1) device resident ----
    G, T = 5, 120
    lats = [torch.empty((T, n), dtype=torch.float64, device=dev) for _ in range(G)]
    outs = [torch.empty((T, n), dtype=torch.float64, device=dev) for _ in range(G)]
    finals = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(G)]
    for m in range(G):
        for r in range(0, T, 24):
            lats[m][r:r + 24] = base * float(scale[m]) * (1.0 + 0.01 * (r // 24))

    def timed(fn, reps=3):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    batched_ms = timed(lambda: plan.route_ensemble_dev(rr.MODE_RAPID, d_q0.data_ptr(), [t.data_ptr() for t in lats], n,
                                                       [t.data_ptr() for t in outs], n, [t.data_ptr() for t in finals], T, 1, stream))
    m_chk = first_basins(down, 60_000)
    checks = []
    for m in (0, G - 1):
        p_out, p_q = subset_parity(down, k, x, m_chk, lats[m][:, :m_chk].cpu().numpy(), outs[m][:, :m_chk].cpu().numpy(),
                                   q_init[:m_chk], finals[m][:m_chk].cpu().numpy())
        checks.append(dict(member=m, parity_reaches=m_chk, parity=p_out, parity_state=p_q))
    keep = outs[G - 1].clone()
    d_q = torch.empty(n, dtype=torch.float64, device=dev)

    def one_by_one():
        for m in range(G):
            d_q.copy_(d_q0)
            plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), lats[m].data_ptr(), n, outs[m].data_ptr(), n, T, 1, stream)
    single_ms = timed(one_by_one)
    same = bool(torch.equal(keep, outs[G - 1]))
    emit(config='C5', stage='ensemble_device_resident', members_per_call=G, reaches=n, steps=T, batched_ms=batched_ms,
         one_call_per_member_ms=single_ms, reach_steps_members_per_s=n * T * G / (batched_ms * 1e-3),
         one_call_per_member_rate=n * T * G / (single_ms * 1e-3), batched_equals_single_bitwise=same, checks=checks)
    del lats, outs, keep
    torch.cuda.empty_cache()
    # ---- (2) host arrays, float32 in / float32 out ----
    G, T = 6, 96
    base_h = base.cpu().numpy()
    h_lat = [rr.pinned_empty((T, n), dtype=np.float32) for _ in range(G)]
    h_out = [rr.pinned_empty((T, n), dtype=np.float32) for _ in range(G)]
    for m in range(G):
        for r in range(0, T, 24):
            h_lat[m][r:r + 24] = (base_h * float(scale[m]) * (1.0 + 0.01 * (r // 24))).astype(np.float32)
    states = np.empty((G, n))
    plan.route_ensemble_host(rr.MODE_RAPID, q_init, h_lat, h_out, 1, q_final=states)          # warm (allocates buffers)
    t = time.perf_counter()
    mean = plan.route_ensemble_host(rr.MODE_RAPID, q_init, h_lat, h_out, 1, q_final=states)
    wall = time.perf_counter() - t
    p_out, p_q = subset_parity(down, k, x, m_chk, h_lat[G - 1][:, :m_chk].astype(np.float64), h_out[G - 1][:, :m_chk].astype(np.float64),
                               q_init[:m_chk], states[G - 1][:m_chk])
    emit(config='C5', stage='ensemble_host_f32', members_per_call=G, reaches=n, steps=T, wall_s=wall,
         reach_steps_members_per_s=n * T * G / wall, h2d_bytes=int(G * T * n * 4), d2h_bytes=int(G * T * n * 4),
         mean_equals_numpy=bool(np.array_equal(mean, np.array(list(states)).mean(axis=0))),
         float32_parity_vs_oracle=p_out, state_parity=p_q,
         full_c5_estimate_s=51 * 360 * n / (n * T * G / wall))
    plan.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c2', 'c3', 'c3_pipeline', 'c4', 'c5']
    for w in which:
        {'c1': c1, 'c2': c2, 'c3': c3, 'c3_pipeline': c3_pipeline, 'c4': c4, 'c5': c5}[w]()
