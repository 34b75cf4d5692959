"""
The flag-free hand-over of small networks (rr_direct.cu, narrow_item) under an adversarial CPU memory model
(tests/sentinel_model.py): any number of warps, random interleaving, relaxed stores that become visible entry by entry in
random order.  Every run must finish (no deadlock), never consume an unwritten entry, and reproduce the strict oracle
bit for bit.
"""
import numpy as np
import pytest

import river_route_b200 as rr
from river_route_b200 import synth
from oracle import oracle
from tests.helpers import network_arrays
from tests.sentinel_model import run_model

NETWORKS = {
    'deep': dict(n=700, n_basins=2, seed=12, depth_bias=0.95),
    'bushy': dict(n=900, n_basins=4, seed=11, depth_bias=0.1),
    'stem': dict(n=500, n_basins=1, seed=13, depth_bias=0.5, main_stem=150),
}


@pytest.mark.parametrize('name', list(NETWORKS))
@pytest.mark.parametrize('n_warps,drain_prob,seed', [(1, 0.5, 0), (2, 0.05, 1), (5, 0.5, 2), (5, 0.9, 3), (64, 0.3, 4), (64, 0.02, 5)])
def test_protocol_is_safe_and_exact_under_any_interleaving(name, n_warps, drain_prob, seed):
    down = synth.forest(**NETWORKS[name])
    n = down.shape[0]
    T, rows_tile = 40, 32                                    # two tiles, the second one short (8 rows: a half group)
    k, x = synth.muskingum_params(n, 1)
    a = network_arrays(down, k, x, 3600, 3600)
    rng = np.random.default_rng(7)
    q0 = rng.uniform(0, 50, n)
    ql = synth.lateral_volumes(T, n, 2)
    plan = rr.Plan(down, renumber='always')
    # RapidMuskingum
    q, ref = q0.copy(), np.zeros((T, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 1)
    out, qs, stats = run_model(plan, True, a['c1'], a['c2'], a['c3'], a['c4_dt'], q0, ql, T, rows_tile, n_warps, seed, drain_prob)
    assert np.array_equal(out, ref) and np.array_equal(qs, q), ('rapid', stats)
    if n_warps > 1 and plan.info['max_block_level'] > 0:
        assert stats['polls'] + stats['refetches'] > 0, 'the run never had to wait: the model did not exercise the protocol'
    # Muskingum (no lateral inflow)
    if seed % 2 == 0:
        q, ref = q0.copy(), np.zeros((T, n))
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q, ref, T, 1)
        out, qs, stats = run_model(plan, False, a['c1'], a['c2'], a['c3'], None, q0, None, T, rows_tile, n_warps, seed + 100, drain_prob)
        assert np.array_equal(out, ref) and np.array_equal(qs, q), ('muskingum', stats)
    plan.close()


@pytest.mark.parametrize('n_warps,drain_prob,seed', [(1, 0.5, 0), (5, 0.05, 1), (64, 0.5, 2)])
def test_progress_flag_protocol_under_the_same_model(n_warps, drain_prob, seed):
    """The hand-over large networks keep on their narrow levels (a release per 16-row group, acquire polls of the upstream
    blocks' counters) in the same adversarial model: after the acquire every entry of the group must be visible."""
    down = synth.forest(**NETWORKS['deep'])
    n = down.shape[0]
    T, rows_tile = 40, 32
    k, x = synth.muskingum_params(n, 1)
    a = network_arrays(down, k, x, 3600, 3600)
    q0 = np.random.default_rng(7).uniform(0, 50, n)
    ql = synth.lateral_volumes(T, n, 2)
    plan = rr.Plan(down, renumber='always')
    q, ref = q0.copy(), np.zeros((T, n))
    oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, 1)
    out, qs, stats = run_model(plan, True, a['c1'], a['c2'], a['c3'], a['c4_dt'], q0, ql, T, rows_tile, n_warps, seed, drain_prob,
                               protocol='flags')
    assert np.array_equal(out, ref) and np.array_equal(qs, q), stats
    plan.close()


def test_the_model_catches_a_broken_protocol(monkeypatch):
    """Sanity of the checker itself: a consumer that trusts the watched word alone (no validation of its own entries) must
    be caught consuming an unwritten entry under out-of-order visibility."""
    import tests.sentinel_model as sm
    src = open(sm.__file__).read()
    broken = src.replace("if accept or not ok[:, k, v].all():", "if False:", 1).replace(
        "                    if ok[:, :, :nv].all():\\n                        break", "                    break", 1)
    assert broken != src
    ns = {}
    exec(compile(broken, 'sentinel_model_broken', 'exec'), ns)
    down = synth.forest(**NETWORKS['deep'])
    n = down.shape[0]
    k, x = synth.muskingum_params(n, 1)
    a = network_arrays(down, k, x, 3600, 3600)
    q0 = np.random.default_rng(7).uniform(0, 50, n)
    ql = synth.lateral_volumes(40, n, 2)
    plan = rr.Plan(down, renumber='always')
    with pytest.raises(AssertionError, match='still shows the pattern'):
        for seed in range(6):
            ns['run_model'](plan, True, a['c1'], a['c2'], a['c3'], a['c4_dt'], q0, ql, 40, 32, 16, seed, 0.6)
    # ... and a progress flag written without the fence (the per-group release turned into a plain store) is caught too
    nofence = src.replace("            for idx in order:                                     # the release makes every earlier store of the warp visible ...\n                apply(bufs[w][idx])\n            bufs[w].clear()",
                          "            if tok[3]:\n                for idx in order:\n                    apply(bufs[w][idx])\n                bufs[w].clear()", 1)
    assert nofence != src
    ns2 = {}
    exec(compile(nofence, 'sentinel_model_nofence', 'exec'), ns2)
    with pytest.raises(AssertionError, match='flag acquired but an entry'):
        for seed in range(6):
            ns2['run_model'](plan, True, a['c1'], a['c2'], a['c3'], a['c4_dt'], q0, ql, 40, 32, 16, seed, 0.2, protocol='flags')
    plan.close()
