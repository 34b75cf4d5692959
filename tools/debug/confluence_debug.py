"""Where does a wide-confluence network differ from the oracle?  (debug helper)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import river_route_b200 as rr
from river_route_b200 import synth
from oracle import oracle
from tests.helpers import network_arrays

max_in = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, T = 20000, 40
rng = np.random.default_rng(max_in)
down = np.full(n, -1, dtype=np.int32)
indeg = np.zeros(n, dtype=np.int64)
for i in range(n - 1):
    if rng.random() < 0.002:
        continue
    for _ in range(8):
        d = int(rng.integers(i + 1, min(n, i + 400)))
        if indeg[d] < max_in:
            down[i] = d
            indeg[d] += 1
            break
k, x = synth.muskingum_params(n, 2)
a = network_arrays(down, k, x, 3600, 3600)
plan = rr.Plan(down, renumber='always', staging='auto')
plan.set_coefficients(a['c1'], a['c2'], a['c3'], a['c4_dt'])
arr = plan.arrays()
q0 = rng.uniform(0, 50, n)
ql = synth.lateral_volumes(T, n, 5)
for mode, name in ((rr.MODE_RAPID, 'rapid'), (rr.MODE_MUSKINGUM, 'musk'), (rr.MODE_MUSKINGUM, 'musk2'), (rr.MODE_RAPID, 'rapid2')):
    q_ref, ref = q0.copy(), np.zeros((T, n))
    if mode == rr.MODE_RAPID:
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_ref, ql, ref, 1)
    else:
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_ref, ref, T, 1)
    q, out = q0.copy(), np.full((T, n), np.nan)
    plan.route_host(mode, q, ql if mode == rr.MODE_RAPID else None, out, 1)
    bad = ~np.isclose(out, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    print(name, 'bad entries', int(bad.sum()), 'nan', int(np.isnan(out).sum()), 'bad reaches', int(bad.any(axis=0).sum()))
    if bad.any():
        cols = np.flatnonzero(bad.any(axis=0))
        inv = arr['inv']
        slot = inv[cols]
        lvl = arr['blk_level'][slot >> 5]
        first_rows = bad[:, cols].argmax(axis=0)
        print(' levels of bad reaches (min, max):', int(lvl.min()), int(lvl.max()), ' indeg:', np.bincount(indeg[cols]))
        order = np.argsort(lvl)[:8]
        for o in order:
            c = cols[o]
            print('  reach', int(c), 'slot', int(slot[o]), 'lane', int(slot[o] & 31), 'level', int(lvl[o]), 'indeg', int(indeg[c]), 'first bad row', int(first_rows[o]),
                  'gpu', out[first_rows[o]:first_rows[o] + 3, c], 'ref', ref[first_rows[o]:first_rows[o] + 3, c])
plan.close()
