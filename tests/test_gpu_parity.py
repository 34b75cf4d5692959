"""
Parity of the CUDA path (through the C ABI of librr_b200.so) with the reference.

  * golden fixtures: outputs of the reference's own numba/scipy code (tests/golden/, oracle/make_golden.py),
    called through drop-in functions with the reference kernels' exact signatures;
  * seeded synthetic networks at sizes the CPU oracle finishes in seconds;
  * size-independent properties at larger sizes (splitting a run in time, superposition of the unclamped state).

Tolerance (BASELINE.json north_star): fp64 discharge within 1e-10 relative, measured as
|gpu - ref| <= tol*|ref| + tol*max_t|ref_reach|; reach ordering / indexing exact.
"""
import numpy as np
import pytest

import river_route_b200 as rr
from river_route_b200 import synth
from oracle import oracle
from tests.conftest import require_cuda
from tests.helpers import network_arrays, parity_error

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True)
def _cuda():
    require_cuda()


# ------------------------------------------------------------------------------------------------------
# golden vectors through the reference kernels' own call signatures
# ------------------------------------------------------------------------------------------------------
def test_rapid_route_golden(route_golden):
    g = route_golden
    q = g['q0'].copy()
    out = np.zeros_like(g['rapid_out'])
    rr.rapid_route(g['csc_indptr'], g['csc_indices'], g['lhs_off'], g['c2'], g['c3'], g['c4_dt'], q, g['ql'], out,
                   int(g['substeps']))
    assert parity_error(out, g['rapid_out']) < TOL
    assert parity_error(q, g['rapid_q']) < TOL
    assert (out >= 0).all()


def test_muskingum_route_golden(route_golden):
    g = route_golden
    q = g['q0'].copy()
    out = np.zeros_like(g['musk_out'])
    rr.muskingum_route(g['csc_indptr'], g['csc_indices'], g['lhs_off'], g['c2'], g['c3'], q, out, int(g['musk_nout']),
                       int(g['musk_nrpo']))
    assert parity_error(out, g['musk_out']) < TOL
    assert parity_error(q, g['musk_q']) < TOL


def test_unit_route_golden(route_golden):
    g = route_golden
    inner, hw = g['inner_idx'], g['hw_idx']
    c1i, c2i, c3i = g['c1'][inner], g['c2'][inner], g['c3'][inner]
    lhs = np.ascontiguousarray(-c1i[g['a_inner_indices']])
    q_ch = g['q0'][inner].copy()
    q_full = q_ch.copy()
    out = np.zeros_like(g['unit_out'])
    rr.unit_route(g['a_inner_indptr'], g['a_inner_indices'], lhs,
                  g['a_inner_indptr'], g['a_inner_indices'], np.ones(g['a_inner_indices'].shape[0]),
                  g['a_hw_indptr'], g['a_hw_indices'], np.ones(g['a_hw_indices'].shape[0]),
                  c1i, c2i, c3i, hw, inner, q_ch, q_full, g['conv'], out, int(g['substeps']))
    q_final = np.empty_like(g['q0'])
    q_final[hw] = g['conv'][-1][hw]
    q_final[inner] = q_full
    assert parity_error(out, g['unit_out']) < TOL
    assert parity_error(q_final, g['unit_q']) < TOL
    assert np.array_equal(out[:, hw], g['conv'][:, hw])  # headwaters: lateral inflow passed through untouched


def test_muskingum_zero_state_gives_exact_zeros():
    """tests/test_muskingum.py:48-73 of the reference."""
    down = synth.forest(4000, 3, seed=2, depth_bias=0.6)
    k, x = synth.muskingum_params(4000, 2)
    a = network_arrays(down, k, x, 3600)
    q = np.zeros(4000)
    out = np.full((6, 4000), 7.0)
    rr.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q, out, 6, 1)
    assert np.all(out == 0) and np.all(q == 0)


# ------------------------------------------------------------------------------------------------------
# seeded synthetic networks vs the CPU oracle
# ------------------------------------------------------------------------------------------------------
def _orders(down):
    yield 'growth', down
    yield 'level', synth.relabel(down, synth.level_sorted_order(down))
    yield 'shuffled', synth.relabel(down, synth.random_topological_order(down, 3))


SYNTH = [
    # n, basins, depth_bias, main_stem, T, dt_runoff, dt_routing, plan options
    (20000, 5, 0.5, 0, 70, 10800, 10800, {}),                    # default dt_routing: many c3 < 0, exercises the clamp
    (20000, 2, 0.9, 0, 40, 10800, 900, {}),                      # 12 substeps
    (15000, 1, 0.5, 2000, 50, 3600, 3600, dict(time_tile=8)),    # deep main stem, short tiles
    (9000, 9, 0.2, 0, 33, 3600, 1800, dict(time_tile=16, tile_stride=1)),
    (9000, 9, 0.2, 0, 33, 3600, 1800, dict(time_tile=4, raw_budget_bytes=1 << 20)),   # tiny exchange budget -> deep rings off
    (15000, 1, 0.5, 2000, 150, 3600, 3600, dict(time_tile=32)),  # deep main stem: narrow levels consume upstream tiles group by group
    (300000, 20, 0.5, 0, 100, 3600, 3600, {}),                   # wide and narrow levels; headwater blocks routed by the staging kernel
    (31, 1, 0.5, 0, 5, 3600, 3600, {}),                          # a single partial block
    (1, 1, 0.5, 0, 3, 3600, 3600, {}),                           # one reach
]


@pytest.mark.parametrize('renumber', ['never', 'always', 'always-registers', 'always-registers-tiled', 'always-tma',
                                      'always-lateral-grouped', 'always-direct', 'always-direct-nohw'])
@pytest.mark.parametrize('n,nbas,bias,stem,T,dt_runoff,dt_routing,opts', SYNTH)
def test_rapid_and_muskingum_vs_oracle(n, nbas, bias, stem, T, dt_runoff, dt_routing, opts, renumber):
    # 'always': level-sorted working order (register-blocked path on tile-major working arrays); '-registers':
    # row-major working arrays; '-tma': bulk-async-copy staged kernel; 'never': params-file order
    opts = dict(opts, renumber=renumber.split('-')[0], staging=renumber.split('-', 1)[1] if '-' in renumber else 'auto')
    base = synth.forest(n, nbas, seed=n % 97, depth_bias=bias, main_stem=stem)
    k, x = synth.muskingum_params(n, 1)
    K = dt_runoff // dt_routing
    rng = np.random.default_rng(n)
    for label, down in _orders(base):
        a = network_arrays(down, k, x, dt_routing, dt_runoff)
        q0 = rng.uniform(0, 50, n)
        ql = synth.lateral_volumes(T, n, 7)
        plan = rr.Plan(down, **opts)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
        q_ref, ref = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, K)
        q, out = q0.copy(), np.full((T, n), np.nan)
        plan.route_host(rr.MODE_RAPID, q, ql, out, K)
        assert parity_error(out, ref) < TOL, (label, 'rapid')
        assert parity_error(q, q_ref) < TOL, (label, 'rapid state')
        assert np.array_equal(out == 0, ref == 0), (label, 'clamp pattern')
        q_ref, ref = q0.copy(), np.zeros((T, n))
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_ref, ref, T, K)
        q, out = q0.copy(), np.full((T, n), np.nan)
        plan.route_host(rr.MODE_MUSKINGUM, q, None, out, K)
        assert parity_error(out, ref) < TOL, (label, 'muskingum')
        assert parity_error(q, q_ref) < TOL, (label, 'muskingum state')
        plan.close()


@pytest.mark.parametrize('renumber', ['never', 'always', 'always-lateral-grouped'])
@pytest.mark.parametrize('n,nbas,bias,stem,T,dt_runoff,dt_routing,opts', SYNTH[:5])
def test_unit_vs_oracle(n, nbas, bias, stem, T, dt_runoff, dt_routing, opts, renumber):
    opts = dict(opts, renumber=renumber.split('-')[0], staging=renumber.split('-', 1)[1] if '-' in renumber else 'auto')
    base = synth.forest(n, nbas, seed=n % 89, depth_bias=bias, main_stem=stem)
    k, x = synth.muskingum_params(n, 2)
    K = dt_runoff // dt_routing
    rng = np.random.default_rng(n + 1)
    for label, down in _orders(base):
        a = network_arrays(down, k, x, dt_routing, dt_runoff)
        q0 = rng.uniform(0, 50, n)
        conv = synth.lateral_volumes(T, n, 9) / dt_runoff
        sp = oracle.unit_split(down.astype(np.int64))
        inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
        c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
        q_ch = q0[inner].copy()
        q_full = q_ch.copy()
        ref = np.zeros((T, n))
        oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                          q_ch, q_full, conv, ref, K)
        q_final = np.empty(n)
        q_final[hw] = conv[-1][hw]
        q_final[inner] = q_full
        plan = rr.Plan(down, **opts)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
        q, out = q0.copy(), np.full((T, n), np.nan)
        plan.route_host(rr.MODE_UNIT, q, conv, out, K)     # router-level semantics (UnitMuskingum._router)
        assert parity_error(out, ref) < TOL, (label, 'unit')
        assert parity_error(q, q_final) < TOL, (label, 'unit state')
        plan.close()


@pytest.mark.parametrize('renumber', ['never', 'always'])
def test_high_indegree_confluences(renumber):
    """Reaches with more upstreams than the kernel keeps in registers (slow-slot path), in- and cross-block."""
    rng = np.random.default_rng(4)
    n = 3000
    down = np.full(n, -1, dtype=np.int32)
    for i in range(n - 1):
        # many reaches drain straight into a few hubs; hubs chain downstream
        hub = ((i // 100) + 1) * 100 - 1
        if i == hub:
            down[i] = min(n - 1, hub + 100)
        elif rng.random() < 0.5:
            down[i] = hub
        else:
            down[i] = min(hub, i + int(rng.integers(1, 40)))
    down[n - 1] = -1
    assert np.bincount(down[down >= 0]).max() > 8
    k, x = synth.muskingum_params(n, 3)
    a = network_arrays(down, k, x, 1800, 3600)
    q0 = rng.uniform(0, 10, n)
    ql = synth.lateral_volumes(25, n, 3)
    plan = rr.Plan(down, renumber=renumber)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    q_ref, ref = q0.copy(), np.zeros((25, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, 2)
    q, out = q0.copy(), np.zeros((25, n))
    plan.route_host(rr.MODE_RAPID, q, ql, out, 2)
    assert parity_error(out, ref) < TOL and parity_error(q, q_ref) < TOL
    # the same network through UnitMuskingum
    sp = oracle.unit_split(down.astype(np.int64))
    inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
    c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
    conv = ql / 3600.0
    q_ch = q0[inner].copy()
    q_full = q_ch.copy()
    ref = np.zeros((25, n))
    oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                      q_ch, q_full, conv, ref, 2)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], None)
    q, out = q0.copy(), np.zeros((25, n))
    plan.route_host(rr.MODE_UNIT, q, conv, out, 2)
    assert parity_error(out, ref) < TOL


# ------------------------------------------------------------------------------------------------------
# size-independent properties at larger sizes
# ------------------------------------------------------------------------------------------------------
def test_time_split_is_bitwise_identical_and_chunked_streaming():
    """
    Routing T steps in one call == routing them in two calls chained through the state
    (reference: tests/test_rapid_muskingum.py:95-143), bit for bit; and a call long enough to be cut into
    several pinned H2D/D2H chunks inside rr_route_host equals the same run done in short calls.
    """
    n, T = 300000, 1000                       # 2.4 MB rows -> the 1 GiB chunk holds 447 rows -> 3 chunks
    down = synth.forest(n, 40, seed=8, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 8)
    a = network_arrays(down, k, x, 3600, 3600)
    ql = synth.lateral_volumes(T, n, 8)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    q1, out1 = np.full(n, 3.0), np.empty((T, n))
    plan.route_host(rr.MODE_RAPID, q1, ql, out1, 1)
    q2, out2 = np.full(n, 3.0), np.empty((T, n))
    for t0 in range(0, T, 190):
        t1 = min(T, t0 + 190)
        plan.route_host(rr.MODE_RAPID, q2, ql[t0:t1], out2[t0:t1], 1)
    assert np.array_equal(out1, out2) and np.array_equal(q1, q2)
    assert (out1 >= 0).all() and np.isfinite(out1).all()
    # oracle spot check on one whole basin (basins are independent)
    basin, nb, _ = rr.label_basins(down)
    b = int(np.argmax(np.bincount(basin) * (np.bincount(basin) < 30000)))
    idx = np.flatnonzero(basin == b)
    sub = synth.relabel(down, np.concatenate([idx, np.setdiff1d(np.arange(n), idx)]))[:idx.size]
    assert sub.max() < idx.size
    sa = network_arrays(sub, k[idx], x[idx], 3600, 3600)
    q_ref, ref = np.full(idx.size, 3.0), np.zeros((200, idx.size))
    oracle.rapid_route(sa['indptr'], sa['indices'], sa['lhs_off'], sa['c2'], sa['c3'], sa['c4_dt'], q_ref,
                       np.ascontiguousarray(ql[:200, idx]), ref, 1)
    assert parity_error(out1[:200, idx], ref) < TOL


def test_superposition_of_unclamped_state():
    """The solve is linear in (state, lateral): the final state (never clamped) must superpose."""
    n, T = 200000, 64
    down = synth.forest(n, 25, seed=9, depth_bias=0.7)
    k, x = synth.muskingum_params(n, 9)
    a = network_arrays(down, k, x, 900, 3600)
    plan = rr.Plan(down)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    rng = np.random.default_rng(9)
    la, lb = synth.lateral_volumes(T, n, 10), synth.lateral_volumes(T, n, 11)
    qa0, qb0 = rng.uniform(0, 20, n), rng.uniform(0, 20, n)

    def run(q0, lat):
        q, out = q0.copy(), np.empty((T, n))
        plan.route_host(rr.MODE_RAPID, q, lat, out, 4)
        return q

    qa, qb, qc = run(qa0, la), run(qb0, lb), run(2.0 * qa0 - 0.5 * qb0, 2.0 * la - 0.5 * lb)
    expect = 2.0 * qa - 0.5 * qb
    assert np.max(np.abs(qc - expect)) <= 1e-9 * np.max(np.abs(expect))


# ------------------------------------------------------------------------------------------------------
# device-pointer API (torch tensors as plain device buffers) and ensembles
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('renumber', ['never', 'always'])
def test_device_api_and_ensemble(renumber):
    import torch
    n, T, M = 50000, 48, 5
    down = synth.forest(n, 6, seed=12, depth_bias=0.6)
    k, x = synth.muskingum_params(n, 12)
    a = network_arrays(down, k, x, 1800, 3600)
    plan = rr.Plan(down, renumber=renumber)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    rng = np.random.default_rng(12)
    q0 = rng.uniform(0, 30, n)
    lats = [synth.lateral_volumes(T, n, 20 + m) for m in range(M)]
    refs, qrefs = [], []
    for m in range(M):
        q, out = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, lats[m], out, 2)
        refs.append(out)
        qrefs.append(q)
    dev = torch.device('cuda:0')
    stream = torch.cuda.current_stream().cuda_stream
    # single-member device call, padded leading dimension
    ld = n + 13
    d_lat = torch.zeros((T, ld), dtype=torch.float64, device=dev)
    d_lat[:, :n] = torch.from_numpy(lats[0]).to(dev)
    d_out = torch.full((T, ld), float('nan'), dtype=torch.float64, device=dev)
    d_q = torch.from_numpy(q0).to(dev)
    plan.route_dev(rr.MODE_RAPID, d_q.data_ptr(), d_lat.data_ptr(), ld, d_out.data_ptr(), ld, T, 2, stream)
    torch.cuda.synchronize()
    assert parity_error(d_out[:, :n].cpu().numpy(), refs[0]) < TOL
    assert parity_error(d_q.cpu().numpy(), qrefs[0]) < TOL
    assert torch.isnan(d_out[:, n:]).all()                      # padding columns are never written
    # ensemble: members routed from the same initial state in one launch (TransformMuskingum.py:121-126)
    d_lats = [torch.from_numpy(l).to(dev) for l in lats]
    d_outs = [torch.empty((T, n), dtype=torch.float64, device=dev) for _ in range(M)]
    d_qf = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(M)]
    d_q0 = torch.from_numpy(q0).to(dev)
    plan.route_ensemble_dev(rr.MODE_RAPID, d_q0.data_ptr(), [t.data_ptr() for t in d_lats], n,
                            [t.data_ptr() for t in d_outs], n, [t.data_ptr() for t in d_qf], T, 2, stream)
    torch.cuda.synchronize()
    for m in range(M):
        assert parity_error(d_outs[m].cpu().numpy(), refs[m]) < TOL, m
        assert parity_error(d_qf[m].cpu().numpy(), qrefs[m]) < TOL, m
    assert np.array_equal(d_q0.cpu().numpy(), q0)               # shared initial state untouched
    # ensemble final state = mean over members in member order (TransformMuskingum.py:145-146)
    mean_ref = np.array(qrefs).mean(axis=0)
    mean_gpu = torch.stack(d_qf).mean(dim=0).cpu().numpy()
    assert parity_error(mean_gpu, mean_ref) < TOL


@pytest.mark.parametrize('max_in', [3, 4, 6])
@pytest.mark.parametrize('staging', ['auto', 'direct'])
def test_confluences_of_many_rivers(max_in, staging):
    """Networks with confluences of three and four reaches run the wide instantiation of the direct kernel (upstream
    sums still in ascending params-file index); beyond four upstreams the plan leaves the direct pipeline and the general
    path takes over."""
    n, T = 20000, 40
    rng = np.random.default_rng(max_in)
    down = np.full(n, -1, dtype=np.int32)
    indeg = np.zeros(n, dtype=np.int64)
    for i in range(n - 1):
        if rng.random() < 0.002:
            continue                                   # an outlet
        for _ in range(8):
            d = int(rng.integers(i + 1, min(n, i + 400)))
            if indeg[d] < max_in:
                down[i] = d
                indeg[d] += 1
                break
    assert indeg.max() == max_in
    k, x = synth.muskingum_params(n, 2)
    a = network_arrays(down, k, x, 3600, 3600)
    plan = rr.Plan(down, renumber='always', staging=staging)
    assert plan.info['max_indegree'] == max_in and plan.info['all_fast'] == int(max_in <= 4)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    q0 = rng.uniform(0, 50, n)
    ql = synth.lateral_volumes(T, n, 5)
    q_ref, ref = q0.copy(), np.zeros((T, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, 1)
    q, out = q0.copy(), np.full((T, n), np.nan)
    plan.route_host(rr.MODE_RAPID, q, ql, out, 1)
    assert parity_error(out, ref) < TOL and parity_error(q, q_ref) < TOL
    q_ref, ref = q0.copy(), np.zeros((T, n))
    oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_ref, ref, T, 1)
    q, out = q0.copy(), np.full((T, n), np.nan)
    plan.route_host(rr.MODE_MUSKINGUM, q, None, out, 1)
    assert parity_error(out, ref) < TOL and parity_error(q, q_ref) < TOL
    plan.close()
