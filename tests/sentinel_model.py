"""
CPU model of the flag-free hand-over of small networks (rr_direct.cu, narrow_item) -- test infrastructure.

compute-sanitizer's racecheck is closed on this pool, so the protocol is exercised here under an adversarial memory
model instead: W "warps" draw (block, tile) tickets in the plan's order and run the item's steps as coroutines that an
outside scheduler interleaves at random; every data store goes into the warp's store buffer and becomes visible
(entry by entry, 8 bytes at a time) in RANDOM order at random later times -- relaxed stores to different addresses are
not ordered -- except that a release (the tile flag) drains the warp's buffer first.  Loads see memory as it is at that
moment.  An item follows narrow_item: optimistic fetch of a 16-entry group of every upstream reach, the watched word (last
entry of the group in a reach of the deepest upstream block), re-fetch of whatever still shows the "not written yet"
pattern, a bounded-spin safety net on the upstream blocks' tile flags; 16 dependent steps; stores; one release per tile.

The model fails if an item ever consumes an entry that still shows the pattern, if the run stops making progress
(deadlock / lost wake-up), or if the result differs from the strict oracle in a single bit (the arithmetic is the
reference's order of operations, without fused multiply-adds).
"""
from __future__ import annotations

import numpy as np

B = 32
GROUP = 16


class Deadlock(AssertionError):
    pass


def run_model(plan, has_lat, c1, c2, c3, c4, q0, lat, T, rows_tile, n_warps, seed, drain_prob, protocol='pattern'):
    """Returns (out [T, n] clamped like stage_out, final state [n], stats).  protocol='flags': the hand-over of large
    networks' narrow levels (direct_item): a block releases done[block] after every 16-row group and a consumer reads a
    group of its upstream reaches only after it has acquired the counters of all its upstream blocks."""
    a = plan.arrays()
    perm, inv = a['perm'], a['inv']
    assert perm is not None, 'the direct pipeline works on level-sorted (renumbered) plans'
    n_user = plan.n
    n = perm.shape[0]
    assert n % B == 0
    nb = n // B

    def spread(v, rows=False):
        w = np.zeros(((v.shape[0], n) if rows else (n,)))
        w[..., inv] = np.asarray(v, dtype=np.float64)
        return w
    c1, c2, c3 = spread(c1), spread(c2), spread(c3)
    c4 = spread(c4) if has_lat else np.zeros(n)
    q_init = spread(q0)
    latw = spread(lat, rows=True) if has_lat else np.zeros((T, n))
    up_ptr, up_idx, dep_ptr, dep_idx, lvl = a['up_ptr'], a['up_idx'], a['dep_ptr'], a['dep_idx'], a['blk_level']
    deg = np.diff(up_ptr).astype(np.int64)
    D = int(deg.max()) if n else 0
    ups = np.full((n, max(D, 1)), -1, dtype=np.int64)            # ascending params-file index per reach (the plan's order)
    for i in range(n):
        ups[i, :deg[i]] = up_idx[up_ptr[i]:up_ptr[i + 1]]
    assert rows_tile % GROUP == 0
    gpt = rows_tile // GROUP
    n_tiles = (T + rows_tile - 1) // rows_tile
    blocks, tiles = plan.schedule(n_tiles, gpt)

    # ---- "device memory" ----
    val = np.zeros((n_tiles, n, rows_tile))
    armed = np.ones((n_tiles, n, rows_tile), dtype=bool)          # fill_sentinel_kernel
    q_state = np.zeros(n)
    done = np.zeros(nb, dtype=np.int64)
    rng = np.random.default_rng(seed)
    stats = {'polls': 0, 'refetches': 0, 'safety_net': 0, 'steps': 0}
    bufs = [[] for _ in range(n_warps)]                           # per warp: ('o', tile, slot, entry, value) | ('q', slot, value)

    def apply(st):
        if st[0] == 'o':
            _, j, slot, e, v = st
            val[j, slot, e] = v
            armed[j, slot, e] = False
        else:
            q_state[st[1]] = st[2]

    def item(w, b, j):
        lanes = np.arange(b * B, b * B + B)
        TT = min(rows_tile, T - j * rows_tile)
        full_want = (j + 1) * gpt
        while done[b] < j * gpt:                                  # own previous tile (acquire)
            yield 'spin'
        q = q_init[lanes].copy() if j == 0 else q_state[lanes].copy()
        yield 'step'
        U_idx = ups[lanes]                                        # [32, D]
        has = U_idx >= 0
        safe_idx = np.where(has, U_idx, 0)
        old = np.zeros((B, max(D, 1)))
        if j == 0:
            old = np.where(has, q_init[safe_idx], 0.0)
        else:
            while True:                                           # last entry of the previous tile of every upstream reach
                set_ = ~armed[j - 1, safe_idx, rows_tile - 1] | ~has
                if set_.all():
                    old = np.where(has, val[j - 1, safe_idx, rows_tile - 1], 0.0)
                    break
                yield 'spin'
        # the watched reach: one upstream reach of the deepest upstream block of the whole warp
        hint = -1
        if has.any():
            lv = np.where(has, lvl[safe_idx // B], -1)
            hint = int(safe_idx.flat[int(np.argmax(lv))])
        up_blocks = np.unique(safe_idx[has] // B)
        for s0 in range(0, TT, GROUP):
            ne = min(GROUP, TT - s0)
            nv = (ne + 3) // 4
            U = np.zeros((B, max(D, 1), GROUP))
            ok = np.ones((B, max(D, 1), 4), dtype=bool)

            def fetch(k, v):
                e0 = s0 + 4 * v
                sl = slice(e0, e0 + 4)
                U[:, k, 4 * v:4 * v + 4] = np.where(has[:, k, None], val[j, safe_idx[:, k], sl], 0.0)
                ok[:, k, v] = ~has[:, k] | ~armed[j, safe_idx[:, k], sl].any(axis=1)
            if protocol == 'flags' and up_blocks.size:
                want = j * gpt + s0 // GROUP + 1
                while (done[up_blocks] < want).any():             # acquire poll of every upstream block's counter
                    stats['polls'] += 1
                    yield 'spin'
            for k in range(D):
                for v in range(nv):
                    fetch(k, v)                                   # optimistic: one "load instruction" per scheduler step
                    yield 'step'
            last = min(s0 + GROUP - 1, TT - 1)
            hv = hint < 0 or not armed[j, hint, last]
            if protocol == 'flags':
                assert ok[:, :, :nv].all(), 'flag acquired but an entry of the group is not visible'
            if not ok[:, :, :nv].all():
                spins, accept = 0, False
                while True:
                    while not hv:
                        spins += 1
                        stats['polls'] += 1
                        if spins % 8 == 0 and (done[up_blocks] >= full_want).all():
                            accept = True                         # safety net: every upstream block released this tile
                            stats['safety_net'] += 1
                            break
                        yield 'spin'
                        hv = not armed[j, hint, last]
                    for k in range(D):
                        for v in range(nv):
                            if accept or not ok[:, k, v].all():
                                fetch(k, v)
                                stats['refetches'] += 1
                                yield 'step'
                    if ok[:, :, :nv].all():
                        break
                    if accept:
                        raise AssertionError('an upstream block released its tile but an entry still shows the pattern')
                    spins += 1
                    if spins % 8 == 0 and (done[up_blocks] >= full_want).all():
                        accept = True
                        stats['safety_net'] += 1
                    yield 'spin'
            assert ok[:, :, :nv].all(), 'an entry that still shows the pattern would be consumed'
            res = np.zeros((B, GROUP))
            for e in range(ne):
                t = j * rows_tile + s0 + e
                r = c3[lanes] * q                                 # _numba_kernels.py:27-28 / :68-69, no contraction
                if has_lat:
                    r = r + c4[lanes] * latw[t, lanes]
                for k in range(D):                                # :29-33 / :70-74 ascending upstream
                    r = np.where(has[:, k], r + c2[lanes] * old[:, k], r)
                for k in range(D):                                # :36-39 / :75-78 (lhs_off = -c1)
                    r = np.where(has[:, k], r + c1[lanes] * U[:, k, e], r)
                q = r
                for k in range(D):
                    old[:, k] = U[:, k, e]
                res[:, e] = r
            yield 'step'
            for v in range(nv):                                   # strong stores: buffered, visible later in any order
                for e in range(4 * v, min(4 * v + 4, ne)):
                    for ln in range(B):
                        bufs[w].append(('o', j, int(lanes[ln]), s0 + e, float(res[ln, e])))
                yield 'step'
            if protocol == 'flags' and s0 + GROUP < TT:
                yield ('release', b, j * gpt + s0 // GROUP + 1, False)   # publish the group: fence, then the counter
        for ln in range(B):
            bufs[w].append(('q', int(lanes[ln]), float(q[ln])))
        yield ('release', b, full_want, True)

    # ---- scheduler ----
    n_items = blocks.shape[0]
    ticket = 0
    warps = [None] * n_warps
    idle_spins = 0
    limit = 400 * n_items * n_warps + 200000
    while True:
        for w in range(n_warps):
            if warps[w] is None and ticket < n_items:
                warps[w] = item(w, int(blocks[ticket]), int(tiles[ticket]))
                ticket += 1
        active = [w for w in range(n_warps) if warps[w] is not None]
        pending = [w for w in range(n_warps) if bufs[w]]
        if not active and not pending:
            break
        if pending and (not active or rng.random() < drain_prob):
            w = pending[int(rng.integers(len(pending)))]
            apply(bufs[w].pop(int(rng.integers(len(bufs[w])))))   # any entry of any warp, in any order
            idle_spins = 0
            continue
        w = active[int(rng.integers(len(active)))]
        try:
            tok = next(warps[w])
        except StopIteration:
            warps[w] = None
            continue
        stats['steps'] += 1
        if tok == 'spin':
            idle_spins += 1
            if idle_spins > limit and not pending:
                raise Deadlock(f'no progress: {len(active)} warps spinning, ticket {ticket} of {n_items}')
        else:
            idle_spins = 0
        if isinstance(tok, tuple) and tok[0] == 'release':
            order = rng.permutation(len(bufs[w]))
            for idx in order:                                     # the release makes every earlier store of the warp visible ...
                apply(bufs[w][idx])
            bufs[w].clear()
            done[tok[1]] = tok[2]                                 # ... before the flag
            if tok[3]:
                warps[w] = None
    for t in range(T):
        assert not armed[t // rows_tile, :, t % rows_tile].any(), 'an entry of the call was never written'
    # stage_out: working tiles -> caller's rows with the reference's clamp
    out = np.zeros((T, n_user))
    real = perm >= 0
    for t in range(T):
        v = val[t // rows_tile, :, t % rows_tile]
        out[t, perm[real]] = np.where(v[real] > 0.0, v[real], 0.0)
    q_final = np.zeros(n_user)
    q_final[perm[real]] = q_state[real]
    return out, q_final, stats
