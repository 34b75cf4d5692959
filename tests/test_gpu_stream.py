"""
GPU parity of the streamed host paths (through the C ABI):

  * rr_route_host_ex      -- the router loop's tail on the device: dt_discharge resample and float32 cast
                             (routers/TransformMuskingum.py:128-146) must be BIT-identical to numpy on the fp64 array;
  * rr_runoff_route_host  -- gathered grid runoff -> weight table [-> unit hydrograph] -> route in one residency,
                             against the CPU oracle chain (oracle.weights_transform -> uh_convolve -> *_route);
  * chunking              -- many short time chunks (RR_STREAM_CHUNK_ROWS) give bit-identical results to one chunk,
                             including the cumulative-runoff difference and the unit-hydrograph carry-over.
"""
import os

import numpy as np
import pytest

import river_route_b200 as rr
from river_route_b200 import synth
from river_route_b200.transforms import Transform
from oracle import oracle
from tests.conftest import require_cuda
from tests.helpers import network_arrays, parity_error

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True)
def _cuda():
    require_cuda()


@pytest.fixture
def chunk_rows():
    def set_rows(rows):
        if rows:
            os.environ['RR_STREAM_CHUNK_ROWS'] = str(rows)
        else:
            os.environ.pop('RR_STREAM_CHUNK_ROWS', None)
    yield set_rows
    os.environ.pop('RR_STREAM_CHUNK_ROWS', None)


def _network(n, nbas, seed, dt_routing, dt_runoff, order='growth'):
    down = synth.forest(n, nbas, seed=seed, depth_bias=0.5)
    if order == 'shuffled':
        down = synth.relabel(down, synth.random_topological_order(down, 3))
    k, x = synth.muskingum_params(n, seed)
    return down, network_arrays(down, k, x, dt_routing, dt_runoff)


def _numpy_tail(q64, k):
    """TransformMuskingum.py:128-146 on the host."""
    if k > 1:
        q64 = q64.reshape(q64.shape[0] // k, k, q64.shape[1]).mean(axis=1)
    return q64.astype(np.float32)


@pytest.mark.parametrize('mode', ['rapid', 'muskingum'])
@pytest.mark.parametrize('k,rows', [(1, 0), (3, 0), (1, 5), (4, 8), (24, 24)])
def test_device_tail_is_bit_identical_to_numpy(mode, k, rows, chunk_rows):
    n, T, K = 6000 + 13, 48, 2
    down, a = _network(n, 4, 5, 1800, 3600, order='shuffled')
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    ql = synth.lateral_volumes(T, n, 3) if mode == 'rapid' else None
    m = rr.MODE_RAPID if mode == 'rapid' else rr.MODE_MUSKINGUM
    q0 = np.random.default_rng(1).uniform(0, 40, n)
    chunk_rows(0)
    q_ref, out64 = q0.copy(), np.empty((T, n))
    plan.route_host(m, q_ref, ql, out64, K)
    chunk_rows(rows)
    for dtype in (np.float32, np.float64):
        q, out = q0.copy(), np.full((T // k, n), np.nan, dtype=dtype)
        plan.route_host(m, q, ql, out, K, resample=k)
        want = _numpy_tail(out64, k) if dtype == np.float32 else (out64.reshape(T // k, k, n).mean(axis=1) if k > 1 else out64)
        assert np.array_equal(out, want), (mode, k, rows, dtype)
        assert np.array_equal(q, q_ref)
    plan.close()


def _weight_case(n, n_points, seed, f32):
    rng = np.random.default_rng(seed)
    per = rng.integers(1, 7, n)
    river_idx = np.repeat(np.arange(n), per)
    point_idx = rng.integers(0, n_points, river_idx.shape[0])
    vals = rng.dirichlet(np.ones(6), n).ravel()[: river_idx.shape[0]]
    indptr, indices, data = oracle.weights_csr(river_idx, point_idx, vals, n, n_points)
    area = rng.uniform(1e5, 5e8, n)
    return indptr, indices, data, area


def _runoff(T, n_points, seed, f32, cumulative):
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.3, 2e-3, (T, n_points)) * (rng.random((T, n_points)) < 0.4)
    x[rng.random((T, n_points)) < 0.001] = np.nan            # missing cells (runoff.py:331-333)
    if cumulative:
        x = np.nancumsum(x, axis=0)
    return x.astype(np.float32) if f32 else x


@pytest.mark.parametrize('f32,cumulative,rows,k', [(True, False, 0, 1), (True, True, 7, 1), (False, True, 16, 2),
                                                   (True, False, 1, 1), (False, False, 8, 4), (True, True, 1, 1)])
def test_runoff_to_discharge_rapid_vs_oracle(f32, cumulative, rows, k, chunk_rows):
    n, n_points, T, K = 9000 + 5, 2500, 40, 1
    down, a = _network(n, 6, 11, 3600, 3600)
    indptr, indices, data, area = _weight_case(n, n_points, 2, f32)
    x = _runoff(T, n_points, 4, f32, cumulative)
    ql = oracle.weights_transform(indptr, indices, data, x, cumulative=cumulative, area=area)
    q0 = np.random.default_rng(2).uniform(0, 30, n)
    q_ref, ref = q0.copy(), np.zeros((T, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, K)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    tf = Transform(indptr, indices, data, n_points, area=area)
    chunk_rows(0)
    q1, one = q0.copy(), np.empty((T, n))
    plan.runoff_route_host(tf, rr.MODE_RAPID, q1, x, one, K, cumulative=cumulative, as_volumes=True)
    assert parity_error(one, ref) < TOL and parity_error(q1, q_ref) < TOL
    chunk_rows(rows)
    q2, out = q0.copy(), np.empty((T // k, n), dtype=np.float32)
    plan.runoff_route_host(tf, rr.MODE_RAPID, q2, x, out, K, cumulative=cumulative, as_volumes=True, resample=k)
    assert np.array_equal(out, _numpy_tail(one, k))          # chunked == single chunk, bit for bit
    assert np.array_equal(q2, q1)
    # the two-call path (weights_transform on the host arrays, then route_host) agrees as well
    from river_route_b200.transforms import weights_transform
    ql_gpu = weights_transform(indptr, indices, data, x, cumulative=cumulative, area=area)
    q3, two = q0.copy(), np.empty((T, n))
    plan.route_host(rr.MODE_RAPID, q3, ql_gpu, two, K)
    assert np.array_equal(two, one)
    tf.close()
    plan.close()


@pytest.mark.parametrize('rows,n_ks', [(0, 9), (5, 9), (3, 20), (1, 4)])
def test_runoff_to_discharge_unit_vs_oracle(rows, n_ks, chunk_rows):
    n, n_points, T, K = 7000 + 3, 1800, 30, 1
    down, a = _network(n, 5, 13, 3600, 3600)
    indptr, indices, data, area = _weight_case(n, n_points, 6, True)
    rng = np.random.default_rng(8)
    ker = rng.uniform(0, 1, (n_ks, n)) * (rng.random((n_ks, n)) < 0.6)
    s0 = rng.uniform(0, 1e-3, (n_ks, n))
    s0[-1] = 0.0
    q0 = rng.uniform(0, 30, n)
    sp = oracle.unit_split(down.astype(np.int64))
    inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
    c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
    tf = Transform(indptr, indices, data, n_points, area=area).set_unit_hydrograph(ker, s0)
    chunk_rows(rows)
    s_ref, state = s0.copy(), q0.copy()
    q_gpu = q0.copy()
    for f in range(2):                                         # two files: UH carry-over and channel state chain
        x = _runoff(T, n_points, 20 + f, True, False)
        depths = oracle.weights_transform(indptr, indices, data, x)          # depths: as_volumes False for UnitMuskingum
        conv = oracle.uh_convolve(depths, ker, s_ref)
        q_ch = state[inner].copy()
        q_full = q_ch.copy()
        ref = np.zeros((T, n))
        oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                          q_ch, q_full, conv, ref, K)
        state = np.empty(n)
        state[hw] = conv[-1][hw]
        state[inner] = q_full
        out = np.empty((T, n))
        plan.runoff_route_host(tf, rr.MODE_UNIT, q_gpu, x, out, K)
        col = np.max(np.abs(conv), axis=0)
        assert parity_error(out, ref, col) < TOL, f
        assert parity_error(q_gpu, state, col) < TOL, f
    col = np.max(np.abs(conv), axis=0) + np.max(np.abs(s0), axis=0)
    assert parity_error(tf.uh_state(), s_ref, col) < TOL
    tf.close()
    plan.close()


def test_argument_errors():
    n = 500
    down, a = _network(n, 2, 1, 3600, 3600)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    indptr, indices, data, area = _weight_case(n, 100, 1, True)
    tf = Transform(indptr, indices, data, 100)
    x = np.zeros((6, 100), dtype=np.float32)
    with pytest.raises(RuntimeError, match='as_volumes needs catchment areas'):
        plan.runoff_route_host(tf, rr.MODE_RAPID, np.zeros(n), x, np.empty((6, n)), 1, as_volumes=True)
    with pytest.raises(RuntimeError, match='needs a unit hydrograph'):
        plan.runoff_route_host(tf, rr.MODE_UNIT, np.zeros(n), x, np.empty((6, n)), 1)
    with pytest.raises(ValueError, match='does not match'):
        plan.runoff_route_host(tf, rr.MODE_RAPID, np.zeros(n), x[:, :50], np.empty((6, n)), 1)
    with pytest.raises(RuntimeError, match='multiple of the output resampling factor'):
        from river_route_b200._lib import lib, check, as_f64p
        import ctypes as C
        out = np.empty((2, n), dtype=np.float32)
        ql = np.zeros((7, n))
        check(lib.rr_route_host_ex(plan._h, rr.MODE_RAPID, as_f64p(np.zeros(n)), None, as_f64p(ql), n,
                                   out.ctypes.data_as(C.c_void_p), n, 7, 1, 1, 3))
    tf.close()
    plan.close()


@pytest.mark.parametrize('rows', [0, 4])
def test_pinned_and_pageable_callers_agree(rows, chunk_rows):
    """Pinned arrays (rr.pinned_empty) are copied in place; ordinary numpy arrays go through the library's pinned
    bounce buffers with threaded host copies.  Same bits either way, also with padded row strides."""
    n, T = 5000 + 7, 22
    down, a = _network(n, 3, 4, 3600, 3600)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    ql = synth.lateral_volumes(T, n, 11)
    q0 = np.random.default_rng(3).uniform(0, 10, n)
    chunk_rows(rows)
    q_a, out_a = q0.copy(), np.empty((T, n))
    plan.route_host(rr.MODE_RAPID, q_a, ql, out_a, 1)
    p_ql, p_out = rr.pinned_empty((T, n)), rr.pinned_empty((T, n), dtype=np.float32)
    p_ql[:] = ql
    q_b = q0.copy()
    plan.route_host(rr.MODE_RAPID, q_b, p_ql, p_out, 1)
    assert np.array_equal(p_out, out_a.astype(np.float32)) and np.array_equal(q_a, q_b)
    wide_in, wide_out = np.full((T, n + 9), np.nan), np.full((T, n + 5), -1.0)
    wide_in[:, :n] = ql
    q_c = q0.copy()
    plan.route_host(rr.MODE_RAPID, q_c, wide_in[:, :n], wide_out[:, :n], 1)
    assert np.array_equal(wide_out[:, :n], out_a) and np.all(wide_out[:, n:] == -1.0) and np.array_equal(q_c, q_a)
    plan.close()


@pytest.mark.parametrize('rows,k', [(0, 1), (3, 2)])
def test_output_subset_columns(rows, k, chunk_rows):
    """rr_plan_set_output_subset: only the chosen river segments are copied back, bit-identical to the columns of the
    full output, for every host streaming call; q_state stays the full final state."""
    n, n_points, T = 4000 + 9, 900, 24
    down, a = _network(n, 3, 17, 3600, 3600)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    indptr, indices, data, area = _weight_case(n, n_points, 5, True)
    tf = Transform(indptr, indices, data, n_points, area=area)
    x = _runoff(T, n_points, 6, True, False)
    ql = synth.lateral_volumes(T, n, 12)
    q0 = np.random.default_rng(5).uniform(0, 10, n)
    idx = np.array([n - 1, 0, 77, 77, 2048, 31, 32], dtype=np.int32)
    chunk_rows(rows)
    q_full, out_full = q0.copy(), np.empty((T // k, n), dtype=np.float32)
    plan.route_host(rr.MODE_RAPID, q_full, ql, out_full, 1, resample=k)
    g_full, gout_full = q0.copy(), np.empty((T // k, n))
    plan.runoff_route_host(tf, rr.MODE_RAPID, g_full, x, gout_full, 1, as_volumes=True, resample=k)
    plan.set_output_subset(idx)
    q_sub, out_sub = q0.copy(), np.full((T // k, idx.shape[0]), np.nan, dtype=np.float32)
    plan.route_host(rr.MODE_RAPID, q_sub, ql, out_sub, 1, resample=k)
    assert np.array_equal(out_sub, out_full[:, idx]) and np.array_equal(q_sub, q_full)
    g_sub, gout_sub = q0.copy(), np.full((T // k, idx.shape[0]), np.nan)
    plan.runoff_route_host(tf, rr.MODE_RAPID, g_sub, x, gout_sub, 1, as_volumes=True, resample=k)
    assert np.array_equal(gout_sub, gout_full[:, idx]) and np.array_equal(g_sub, g_full)
    with pytest.raises(ValueError, match='shape'):
        plan.route_host(rr.MODE_RAPID, q0.copy(), ql, np.empty((T // k, n), dtype=np.float32), 1, resample=k)
    with pytest.raises(RuntimeError, match='outside the network'):
        plan.set_output_subset([n])
    plan.set_output_subset(None)
    again = np.empty((T // k, n), dtype=np.float32)
    plan.route_host(rr.MODE_RAPID, q0.copy(), ql, again, 1, resample=k)
    assert np.array_equal(again, out_full)
    tf.close()
    plan.close()


@pytest.mark.parametrize('rows', [0, 5])
def test_direct_exchange_staging_matches_default(rows, chunk_rows):
    """staging='direct' (the working discharge array doubles as the exchange buffer) gives the same bits as the default
    exchange rings, also when a call is continued chunk by chunk from the running state."""
    n, T = 9000 + 3, 37
    down, a = _network(n, 4, 23, 3600, 3600)
    ql = synth.lateral_volumes(T, n, 8)
    q0 = np.random.default_rng(9).uniform(0, 20, n)
    res = {}
    chunk_rows(rows)
    for staging in ('registers-tiled', 'direct'):
        plan = rr.Plan(down, renumber='always', staging=staging)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
        q, out = q0.copy(), np.empty((T, n))
        plan.route_host(rr.MODE_RAPID, q, ql, out, 1)
        qm, outm = q0.copy(), np.empty((T, n))
        plan.route_host(rr.MODE_MUSKINGUM, qm, None, outm, 1)
        res[staging] = (q, out, qm, outm)
        plan.close()
    for x, y in zip(res['registers-tiled'], res['direct']):
        assert np.array_equal(x, y)
    assert (res['direct'][1] >= 0).all() and (res['direct'][1] == 0).any()      # clamp applied by the permutation


@pytest.mark.parametrize('staging', ['auto', 'direct', 'registers-tiled'])
@pytest.mark.parametrize('rows', [0, 16, 5])
def test_float32_lateral_inflows_cross_pcie_as_stored(staging, rows, chunk_rows):
    """rr_route_host_typed: a float32 qlateral array gives the very bits of routing its astype(float64) copy
    (TransformMuskingum.py:36), whether the direct pipeline upcasts while staging or another path upcasts first; the
    float32 output / output subset written straight from the working tiles equals the two-pass result."""
    n, T = 40000 + 7, 56
    down, a = _network(n, 6, 13, 3600, 3600)
    plan = rr.Plan(down, renumber='always', staging=staging)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    ql32 = synth.lateral_volumes(T, n, 8).astype(np.float32)
    q0 = np.random.default_rng(4).uniform(0, 40, n)
    chunk_rows(0)
    q_ref, ref = q0.copy(), np.empty((T, n))
    plan.route_host(rr.MODE_RAPID, q_ref, ql32.astype(np.float64), ref, 1)
    chunk_rows(rows)
    q, out = q0.copy(), np.full((T, n), np.nan)
    plan.route_host(rr.MODE_RAPID, q, ql32, out, 1)
    assert np.array_equal(out, ref) and np.array_equal(q, q_ref)
    q, out32 = q0.copy(), np.full((T, n), np.nan, dtype=np.float32)
    plan.route_host(rr.MODE_RAPID, q, ql32, out32, 1)
    assert np.array_equal(out32, ref.astype(np.float32)) and np.array_equal(q, q_ref)
    pick = np.array([n - 1, 0, 777, 0, 31999], dtype=np.int32)
    plan.set_output_subset(pick)
    q, sub = q0.copy(), np.full((T, pick.shape[0]), np.nan, dtype=np.float32)
    plan.route_host(rr.MODE_RAPID, q, ql32, sub, 1)
    assert np.array_equal(sub, ref.astype(np.float32)[:, pick]) and np.array_equal(q, q_ref)
    plan.set_output_subset(None)
    # and against the CPU oracle
    q_o, ref_o = q0.copy(), np.zeros((T, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_o, ql32.astype(np.float64), ref_o, 1)
    assert parity_error(ref, ref_o) < TOL and parity_error(q_ref, q_o) < TOL
    plan.close()


@pytest.mark.parametrize('staging,K', [('auto', 1), ('direct', 1), ('registers-tiled', 1), ('auto', 2)])
@pytest.mark.parametrize('f32,rows', [(False, 0), (True, 16)])
def test_ensemble_host_members_match_oracle(staging, K, f32, rows, chunk_rows):
    """rr_route_ensemble_host: every member equals its own single-member oracle run (1e-10), the float32 rows equal the
    cast of the fp64 result, time slabs chain per-member states bit-identically to one call, and the mean state is
    numpy's member-order mean bit for bit (TransformMuskingum.py:121-126, :145-146)."""
    n, T, M = 25000 + 3, 48, 5
    down, a = _network(n, 5, 17, 3600 // K, 3600)
    plan = rr.Plan(down, renumber='always', staging=staging)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    base = synth.lateral_volumes(T, n, 9)
    lats = [base * s for s in np.random.default_rng(1).lognormal(0, 0.3, M)]
    if f32:
        lats = [x.astype(np.float32) for x in lats]
    q0 = np.random.default_rng(2).uniform(0, 40, n)
    chunk_rows(rows)
    outs = [np.full((T, n), np.nan, dtype=np.float32) for _ in range(M)]
    finals = np.empty((M, n))
    mean = plan.route_ensemble_host(rr.MODE_RAPID, q0, lats, outs, K, q_final=finals)
    refs, qrefs = [], []
    for m in range(M):
        q, ref = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, lats[m].astype(np.float64), ref, K)
        refs.append(ref)
        qrefs.append(q)
        assert parity_error(finals[m], q) < TOL, m
        assert np.allclose(outs[m], ref.astype(np.float32), rtol=3e-7, atol=1e-30), m
        # bit-level: the member alone through the single-member call
        q1, o1 = q0.copy(), np.empty((T, n), dtype=np.float32)
        plan.route_host(rr.MODE_RAPID, q1, lats[m], o1, K)
        assert np.array_equal(o1, outs[m]) and np.array_equal(q1, finals[m]), m
    assert np.array_equal(mean, np.array(list(finals)).mean(axis=0))
    # two time slabs through per-member initial states == one call
    outs_a = [np.empty((32, n), dtype=np.float32) for _ in range(M)]
    outs_b = [np.empty((T - 32, n), dtype=np.float32) for _ in range(M)]
    st = np.empty((M, n))
    plan.route_ensemble_host(rr.MODE_RAPID, q0, [x[:32] for x in lats], outs_a, K, q_final=st)
    st2 = np.empty((M, n))
    mean2 = plan.route_ensemble_host(rr.MODE_RAPID, np.ascontiguousarray(st), [x[32:] for x in lats], outs_b, K, q_final=st2)
    for m in range(M):
        assert np.array_equal(np.vstack([outs_a[m], outs_b[m]]), outs[m]), m
    assert np.array_equal(st2, finals) and np.array_equal(mean2, mean)
    # Muskingum members (no lateral): all members identical to one run
    outs_m = [np.empty((T, n)) for _ in range(2)]
    plan.route_ensemble_host(rr.MODE_MUSKINGUM, q0, None, outs_m, K)
    q1, o1 = q0.copy(), np.empty((T, n))
    plan.route_host(rr.MODE_MUSKINGUM, q1, None, o1, K)
    assert np.array_equal(outs_m[0], o1) and np.array_equal(outs_m[1], o1)
    plan.close()
