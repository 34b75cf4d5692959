"""
Host logic of the product (no GPU): the C-ABI library loads and exports every symbol of include/rr_b200.h,
topology checks raise the reference's errors, the plan's data structures are consistent, the ticket order is
a linear extension of every dependency, and executing the plan's data flow on the CPU (tests/emulator.py)
reproduces the strict oracle BIT FOR BIT -- i.e. the summation order is the reference's.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import river_route_b200 as rr
from river_route_b200 import _lib, synth
from oracle import oracle
from tests.conftest import ROOT, load_golden
from tests.emulator import emulate
from tests.helpers import network_arrays


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'rr_b200.h')).read()
    declared = set(re.findall(r'\b(rr_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/rr_b200.h but not exported'
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.rr_version() >= 100


def test_no_cpu_fallback_without_device():
    if rr.cuda_available():
        pytest.skip('a CUDA device is present')
    down = synth.forest(100, 1, seed=3)
    plan = rr.Plan(down)
    k, x = synth.muskingum_params(100)
    a = network_arrays(down, k, x, 3600, 3600)
    plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
    q = np.zeros(100)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        plan.route_host(rr.MODE_RAPID, q, np.zeros((4, 100)), np.zeros((4, 100)), 1)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        rr.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, np.zeros((4, 100)),
                       np.zeros((4, 100)), 1)


def test_topology_errors_match_reference():
    g = load_golden('tools.npz')
    down = rr.downstream_index(g['ids9'], g['ds9'])
    assert np.array_equal(down, oracle.downstream_index(g['ids9'], g['ds9']))
    with pytest.raises(ValueError) as e:                                   # tests/test_tools.py:48-53
        rr.downstream_index(np.array([10, 20, 30]), np.array([20, -1, 10]))
    assert str(e.value) == str(g['err_unsorted'])
    with pytest.raises(ValueError, match='downstream IDs not in river_id column'):   # Muskingum.py:161-166
        rr.downstream_index(np.array([10, 20]), np.array([-1, 999]))
    with pytest.raises(ValueError, match='Unknown downstream_river_id: 0'):          # id 0
        rr.downstream_index(np.array([10, 20]), np.array([0, -1]))
    with pytest.raises(ValueError, match='duplicate river IDs'):                     # Muskingum.py:153-154
        rr.downstream_index(np.array([10, 10, 30]), np.array([30, 30, -1]))
    with pytest.raises(ValueError, match='topologically sorted'):
        rr.Plan(np.array([1, 0, -1], dtype=np.int32))


@pytest.mark.parametrize('seed', [0, 1])
def test_downstream_index_matches_oracle_on_random_ids(seed):
    rng = np.random.default_rng(seed)
    down = synth.forest(5000, 7, seed=seed, depth_bias=0.6)
    ids = rng.permutation(50000)[:5000].astype(np.int64) + 1
    ds = np.where(down >= 0, ids[np.where(down >= 0, down, 0)], -1)
    assert np.array_equal(rr.downstream_index(ids, ds), down)
    assert np.array_equal(oracle.downstream_index(ids, ds), down)


def test_basin_labels_and_packing():
    down = synth.forest(20000, 37, seed=5, depth_bias=0.5)
    basin, nb, part = rr.label_basins(down, 4)
    assert nb == 37 == int((down < 0).sum())
    # every reach carries its outlet's label, outlets numbered ascending
    outlets = np.flatnonzero(down < 0)
    assert np.array_equal(basin[outlets], np.arange(nb))
    has = down >= 0
    assert np.array_equal(basin[has], basin[down[has]])
    # basins are never cut, parts are balanced as LPT guarantees (<= 4/3 OPT; here: within one basin of the mean)
    assert all(len(set(part[basin == b])) == 1 for b in range(nb))
    load = np.bincount(part, minlength=4)
    assert load.max() - load.min() <= np.bincount(basin).max()


NETWORKS = {
    'bushy': dict(n=1500, n_basins=5, seed=11, depth_bias=0.1),
    'deep': dict(n=1200, n_basins=2, seed=12, depth_bias=0.95),
    'stem': dict(n=900, n_basins=1, seed=13, depth_bias=0.5, main_stem=300),
    'tiny': dict(n=7, n_basins=2, seed=14, depth_bias=0.5),
    'ragged': dict(n=1029, n_basins=3, seed=15, depth_bias=0.7),
}


def _variants(name):
    down = synth.forest(**NETWORKS[name])
    yield 'growth', down
    yield 'level', synth.relabel(down, synth.level_sorted_order(down))
    yield 'shuffled', synth.relabel(down, synth.random_topological_order(down, 3))


@pytest.mark.parametrize('name', list(NETWORKS))
@pytest.mark.parametrize('renumber', ['never', 'always'])
def test_plan_structures(name, renumber):
    for label, user_down in _variants(name):
        plan = rr.Plan(user_down, renumber=renumber)
        a, inf = plan.arrays(), plan.info
        n = user_down.shape[0]
        down = a['down']
        assert inf['renumbered'] == (renumber == 'always') and inf['reach_depth'] == synth.depth(user_down)
        perm = a['perm'] if a['perm'] is not None else np.arange(n)
        real = perm >= 0                                                      # renumbered plans pad levels to whole blocks
        assert np.array_equal(np.sort(perm[real]), np.arange(n)) and perm.shape[0] == inf['n_work']
        assert np.all(down[down >= 0] > np.flatnonzero(down >= 0))          # working order is topological too
        assert np.all(down[~real] < 0)                                        # padding slots are isolated
        # upstream-CSR lists each reach's upstreams in ascending USER index (the reference's summation order)
        for i in range(perm.shape[0]):
            ups = a['up_idx'][a['up_ptr'][i]:a['up_ptr'][i + 1]]
            want = np.flatnonzero(user_down == perm[i]) if real[i] else np.zeros(0, dtype=np.int64)
            assert np.array_equal(perm[ups], want), label
        if renumber == 'always':
            # one level per block: no in-block edges, every block on the register-blocked path, DAG depth = network depth
            lv = synth.levels(user_down)
            for b in range(inf['n_blocks']):
                members = perm[b * 32:(b + 1) * 32]
                assert len(set(lv[members[members >= 0]].tolist())) <= 1, label
            assert inf['n_internal_edges'] == 0 and inf['all_fast'] == 1 and inf['max_skew'] == 0
            assert inf['max_block_level'] == inf['reach_depth'] - 1
            assert inf['n_work'] - n < 32 * inf['reach_depth'] and inf['n_headwaters'] == int((lv == 0).sum())
        n = perm.shape[0]
        blk = np.arange(n) // 32
        has = down >= 0
        internal = has & (blk == np.where(has, down, 0) // 32)
        # in-block edges: the upstream lane is exactly one systolic step ahead
        assert np.all(a['skew'][np.flatnonzero(internal)].astype(int) + 1 == a['skew'][down[internal]]), label
        # exported series exist exactly for reaches whose downstream is in another block
        assert np.array_equal(a['export_id'] >= 0, has & ~internal), label
        assert inf['n_export'] == int((has & ~internal).sum()) and inf['n_internal_edges'] == int(internal.sum())
        # block levels respect every cross-block edge; spans are positive
        ext = np.flatnonzero(has & ~internal)
        assert np.all(a['blk_level'][blk[ext]] < a['blk_level'][down[ext] // 32]), label
        assert np.array_equal(a['exp_span'][a['export_id'][ext]], a['blk_level'][down[ext] // 32] - a['blk_level'][blk[ext]])


@pytest.mark.parametrize('delta', [1, 3, 1000])
@pytest.mark.parametrize('renumber', ['never', 'always'])
def test_ticket_order_is_a_linear_extension(delta, renumber):
    plan = rr.Plan(synth.forest(**NETWORKS['deep']), renumber=renumber)
    a = plan.arrays()
    down = a['down']
    n_tiles = 5
    blocks, tiles = plan.schedule(n_tiles, delta)
    nb = plan.info['n_blocks']
    pos = np.full((nb, n_tiles), -1, dtype=np.int64)
    pos[blocks, tiles] = np.arange(blocks.shape[0])
    assert (pos >= 0).all() and blocks.shape[0] == nb * n_tiles      # every item exactly once
    assert np.all(pos[:, 1:] > pos[:, :-1])                           # own previous tile first
    for b in range(nb):
        for ub in a['dep_idx'][a['dep_ptr'][b]:a['dep_ptr'][b + 1]]:
            assert np.all(pos[ub] < pos[b])                           # upstream block, same tile
    ring = np.minimum(a['exp_span'] // delta + 1, n_tiles)
    for i in np.flatnonzero(a['export_id'] >= 0):
        r = ring[a['export_id'][i]]
        for j in range(r, n_tiles):                                   # ring reuse waits on an EARLIER ticket
            assert pos[down[i] // 32, j - r] < pos[i // 32, j]


CASES = [  # network, variant-independent numerics: (T, K, tile_substeps, delta)
    ('bushy', 9, 1, 8, 1), ('deep', 7, 2, 8, 2), ('stem', 6, 1, 4, 1), ('tiny', 5, 3, 32, 1), ('ragged', 11, 1, 32, 4),
]


@pytest.mark.parametrize('name,T,K,tile,delta', CASES)
@pytest.mark.parametrize('renumber', ['never', 'always'])
def test_emulated_dataflow_is_bit_exact(name, T, K, tile, delta, renumber):
    for label, down in _variants(name):
        n = down.shape[0]
        k, x = synth.muskingum_params(n, 1)
        dt_runoff = 10800
        a = network_arrays(down, k, x, dt_runoff // K, dt_runoff)
        rng = np.random.default_rng(5)
        q0 = rng.uniform(0, 50, n)
        ql = synth.lateral_volumes(T, n, 2)
        plan = rr.Plan(down, renumber=renumber)
        # RapidMuskingum
        q, ref = q0.copy(), np.zeros((T, n))
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q, ql, ref, K)
        out, qs, _ = emulate(plan, rr.MODE_RAPID, a['c1'], a['c2'], a['c3'], a['c4_dt'], q0, ql, T, K, tile, delta)
        assert np.array_equal(out, ref) and np.array_equal(qs, q), (label, 'rapid')
        # Muskingum
        q, ref = q0.copy(), np.zeros((T, n))
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q, ref, T, K)
        out, qs, _ = emulate(plan, rr.MODE_MUSKINGUM, a['c1'], a['c2'], a['c3'], None, q0, None, T, K, tile, delta)
        assert np.array_equal(out, ref) and np.array_equal(qs, q), (label, 'muskingum')
        # UnitMuskingum
        sp = oracle.unit_split(down.astype(np.int64))
        inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
        c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
        conv = ql / dt_runoff
        q_ch = q0[inner].copy()
        q_full = q_ch.copy()
        ref = np.zeros((T, n))
        oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                          q_ch, q_full, conv, ref, K)
        q_final = np.empty(n)
        q_final[hw] = conv[-1][hw]
        q_final[inner] = q_full
        out, qs, _ = emulate(plan, rr.MODE_UNIT, a['c1'], a['c2'], a['c3'], None, q0, conv, T, K, tile, delta)
        assert np.array_equal(out, ref) and np.array_equal(qs, q_final), (label, 'unit')
