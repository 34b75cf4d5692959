"""
Hardware stress test of the flag protocol (compute-sanitizer's racecheck is not available on this pool): the wavefront
kernels synchronise through per-block progress counters (st.release / ld.acquire) and take tickets in an order that must
make ANY number of persistent CTAs and ANY relative timing safe.  Here the same calls are repeated with the grid capped
at 1 .. max CTAs (RR_GRID_CTAS), tile lengths 16 .. 64, and pseudo-random delays of up to 16 us injected around the flag
operations (RR_JITTER), on a deep narrow network (every level consumes its upstream group by group) and a wide one.
Every run must reproduce the first run bit for bit, and that run must match the CPU oracle.  Round 2: small networks hand
results over without flags (tiles armed with a signalling-NaN pattern, rr_direct.cu narrow_item); both protocols run here
on both networks and must give the same bits.
"""
import os

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
    yield
    os.environ.pop('RR_GRID_CTAS', None)
    os.environ.pop('RR_JITTER', None)
    os.environ.pop('RR_SENTINEL', None)


NETS = {
    'deep': dict(n=40000, n_basins=2, seed=5, depth_bias=0.6, main_stem=1500),
    'wide': dict(n=150000, n_basins=40, seed=6, depth_bias=0.3),
}


@pytest.mark.parametrize('net', list(NETS))
@pytest.mark.parametrize('mode,K,staging', [('rapid', 1, 'auto'), ('rapid', 1, 'direct'), ('muskingum', 1, 'auto'),
                                             ('rapid', 3, 'auto'), ('unit', 1, 'auto')])
def test_any_grid_size_and_timing_gives_the_same_bits(net, mode, K, staging):
    down = synth.forest(**NETS[net])
    n = down.shape[0]
    T = 80 if net == 'deep' else 48
    k, x = synth.muskingum_params(n, 3)
    a = network_arrays(down, k, x, 3600 // K, 3600)
    ql = synth.lateral_volumes(T, n, 4) * (1.0 if mode != 'unit' else 1e-7)
    q0 = np.random.default_rng(9).uniform(0, 30, n)
    m = {'rapid': rr.MODE_RAPID, 'muskingum': rr.MODE_MUSKINGUM, 'unit': rr.MODE_UNIT}[mode]
    first = None
    runs = 0
    for tile in (0, 16, 32, 64):
        # both hand-over protocols of the narrow levels on both networks: per-group progress flags (RR_SENTINEL=0, the
        # default of networks above 4096 blocks) and the armed-tile pattern (RR_SENTINEL=1, the default below); same bits
        if tile == 16:
            os.environ['RR_SENTINEL'] = '0'
        elif tile == 32:
            os.environ['RR_SENTINEL'] = '1'
        else:
            os.environ.pop('RR_SENTINEL', None)
        plan = rr.Plan(down, renumber='always', staging=staging, time_tile=tile)
        plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'] if mode == 'rapid' else None)
        for ctas in (0, 1, 2, 3, 17, 148, 295):
            for jitter in ((0, 10, 14) if ctas in (0, 3) else (0, 12)):
                if ctas:
                    os.environ['RR_GRID_CTAS'] = str(ctas)
                else:
                    os.environ.pop('RR_GRID_CTAS', None)
                os.environ['RR_JITTER'] = str(jitter)
                reps = 6 if (jitter == 0 and ctas == 0) else 1
                for _ in range(reps):
                    q, out = q0.copy(), np.empty((T, n))
                    plan.route_host(m, q, None if mode == 'muskingum' else ql, out, K)
                    runs += 1
                    if first is None:
                        first = (out, q)
                    else:
                        assert np.array_equal(out, first[0]) and np.array_equal(q, first[1]), (tile, ctas, jitter)
        plan.close()
    assert runs > 50
    out, q = first
    q_ref, ref = q0.copy(), np.zeros((T, n))
    if mode == 'rapid':
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, K)
    elif mode == 'muskingum':
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_ref, ref, T, K)
    else:
        sp = oracle.unit_split(down.astype(np.int64))
        inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
        c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
        q_ch = q0[inner].copy()
        q_fu = q_ch.copy()
        oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                          q_ch, q_fu, ql, ref, K)
        q_ref[hw] = ql[-1][hw]
        q_ref[inner] = q_fu
    assert parity_error(out, ref) < TOL and parity_error(q, q_ref) < TOL
