"""
CPU emulation of the wavefront kernel's *schedule and data flow* (test infrastructure).

It executes the plan exactly as rr_route.cu does -- tickets in schedule order, one 32-lane
systolic item per ticket, in-block values taken from the neighbour lane's previous step
("shuffles"), cross-block values from per-reach exchange rings guarded by done[] counters --
but with plain non-fused fp64 scalar arithmetic.  Because the summation order is the
reference's, the result must equal the strict (-ffp-contract=off) oracle bit for bit, which
pins the plan data structures, the skews, the ring sizing and the ticket order on the CPU.
Every dependency the kernel would spin on is asserted to be already satisfied, which proves
the ticket order is a linear extension (no deadlock).
"""
from __future__ import annotations

import numpy as np

HW_BIT = 0x40000000
B = 32


def emulate(plan, mode, c1, c2, c3, c4, q0, lat, T, K, tile_substeps=32, delta=1, qfull0=None, router_level=True):
    """Returns (out [T, n], q_state_final [n], q_full_final or None)."""
    UNIT = mode == 2
    HAS_LAT = mode != 0
    a = plan.arrays()
    down = a['down']
    n = plan.n
    perm = a['perm']
    if perm is not None:
        # renumbered plan: run in the working order (every level padded to whole blocks; padding slots are isolated
        # dummy reaches with zero coefficients), hand results back in the caller's order
        inv = a['inv']
        n = perm.shape[0]

        def spread(v, rows=False):
            if v is None:
                return None
            v = np.asarray(v, dtype=np.float64)
            w = np.zeros(((v.shape[0], n) if rows else (n,)))
            w[..., inv] = v
            return w
        c1, c2, c3, c4 = spread(c1), spread(c2), spread(c3), spread(c4)
        q0 = spread(q0)
        lat = spread(lat, rows=True)
        qfull0 = spread(qfull0)
    nb = (n + B - 1) // B
    rows_tile = max(1, min(T, tile_substeps // K))
    n_tiles = (T + rows_tile - 1) // rows_tile
    pitch = ((rows_tile * K + 2 + 3) // 4) * 4
    blocks, tiles = plan.schedule(n_tiles, delta)
    ring = np.minimum(a['exp_span'] // delta + 1, n_tiles)
    raw = [np.full((int(r), pitch), np.nan) for r in ring]
    done = np.zeros(nb, dtype=np.int64)
    q_state = np.array(q0, dtype=np.float64)
    q_full = q_state.copy() if qfull0 is None else np.array(qfull0, dtype=np.float64)
    out = np.full((T, n), np.nan)
    up_ptr, up_idx, slot_src, skew, export_id = a['up_ptr'], a['up_idx'], a['slot_src'], a['skew'], a['export_id']
    dep_ptr, dep_idx = a['dep_ptr'], a['dep_idx']
    inv_k = 1.0 / K
    seen = set()
    for b, j in zip(blocks.tolist(), tiles.tolist()):
        assert (b, j) not in seen
        seen.add((b, j))
        # --- the waits of the kernel must already be satisfied in ticket order ---
        assert done[b] == j, 'own previous tile not finished'
        for ub in dep_idx[dep_ptr[b]:dep_ptr[b + 1]]:
            assert done[ub] >= j + 1, 'upstream block not finished for this tile'
        lanes = [i for i in range(b * B, min(n, b * B + B))]
        for i in lanes:
            e = export_id[i]
            if e >= 0 and j >= ring[e]:
                assert done[down[i] // B] >= j - ring[e] + 1, 'exchange ring would be overwritten too early'
        t0 = j * rows_tile
        rows = min(rows_tile, T - t0)
        TT = rows * K
        max_skew = max(int(skew[i]) for i in lanes)
        first = (j == 0) and router_level
        qcur = {i: q_state[i] for i in lanes}
        qprev = dict(qcur)
        qf_cur = {i: (qcur[i] if (first or not UNIT) else q_full[i]) for i in lanes}
        qf_prev = dict(qf_cur)
        for i in lanes:
            e = export_id[i]
            if e >= 0:
                row = raw[e][j % ring[e]]
                row[:] = np.nan
                row[0] = qcur[i]
                if UNIT:
                    row[pitch - 1] = qf_cur[i]
        acc = {i: 0.0 for i in lanes}
        sub = {i: 0 for i in lanes}
        row_i = {i: 0 for i in lanes}
        base = {i: 0.0 for i in lanes}
        ql = {i: 0.0 for i in lanes}
        lu = {}
        lu_old = {}
        for sig in range(TT + max_skew):
            snap_cur, snap_prev, snap_fprev = dict(qcur), dict(qprev), dict(qf_prev)  # the "shuffle" view
            for i in lanes:
                s = sig - int(skew[i])
                if not (0 <= s < TT):
                    continue
                deg = up_ptr[i + 1] - up_ptr[i]
                srcs = slot_src[up_ptr[i]:up_ptr[i + 1]].tolist()
                ups = up_idx[up_ptr[i]:up_ptr[i + 1]].tolist()

                def is_hw(sk):
                    return bool(sk & HW_BIT) if sk >= 0 else bool(((-sk - 1) >> 6) & 1)

                def ext_row(sk):
                    e = sk & ~HW_BIT
                    return raw[e][j % ring[e]]

                t = t0 + row_i[i]
                if sub[i] == 0:
                    if HAS_LAT:
                        ql[i] = lat[t, i]
                    if UNIT:
                        a_in = 0.0
                        a_hw = 0.0
                        for k in range(deg):
                            lu_old[(i, k)] = lu.get((i, k), 0.0)
                            lu[(i, k)] = lat[t, ups[k]]
                            if is_hw(srcs[k]):
                                a_hw += lu[(i, k)]
                            else:
                                a_in += lu[(i, k)]
                        base[i] = c1[i] * (a_in + a_hw) + c2[i] * a_hw
                if UNIT:
                    r = base[i] + c3[i] * qcur[i]
                elif HAS_LAT:
                    r = c3[i] * qcur[i] + c4[i] * ql[i]
                else:
                    r = c3[i] * qcur[i]
                vo, vn = [], []
                for k in range(deg):
                    sk = srcs[k]
                    if sk < 0:
                        ln = b * B + ((-sk - 1) & 31)
                        assert int(skew[ln]) == int(skew[i]) - 1, 'in-block upstream must lead by exactly one step'
                        vo.append(snap_fprev[ln] if UNIT else snap_prev[ln])
                        vn.append(snap_cur[ln])
                    else:
                        rr_ = ext_row(sk)
                        if UNIT:
                            if is_hw(sk):
                                vo.append(0.0)
                                vn.append(0.0)
                                continue
                            if s == 0:
                                vo.append(rr_[pitch - 1])
                            else:
                                vo.append(rr_[s] + (lu_old[(i, k)] if sub[i] == 0 else lu[(i, k)]))
                        else:
                            vo.append(rr_[s])
                        vn.append(rr_[s + 1])
                        assert not np.isnan(vo[-1]) and not np.isnan(vn[-1]), 'read of an unwritten exchange entry'
                for k in range(deg):
                    if not (UNIT and is_hw(srcs[k])):
                        r = r + c2[i] * vo[k]
                for k in range(deg):
                    if not (UNIT and is_hw(srcs[k])):
                        r = r + c1[i] * vn[k]
                inner = (not UNIT) or deg > 0
                if inner:
                    qprev[i] = qcur[i]
                    qcur[i] = r
                    if UNIT:
                        qf_prev[i] = qf_cur[i]
                        qf_cur[i] = r + ql[i]
                        acc[i] += qf_cur[i]
                    else:
                        acc[i] += r
                    e = export_id[i]
                    if e >= 0:
                        raw[e][j % ring[e]][1 + s] = r
                sub[i] += 1
                if sub[i] == K:
                    if UNIT and not inner:
                        v = ql[i]
                    else:
                        v = acc[i] * inv_k
                        v = v if v > 0.0 else 0.0
                    out[t, i] = v
                    acc[i] = 0.0
                    sub[i] = 0
                    row_i[i] += 1
        last = j == n_tiles - 1
        for i in lanes:
            deg = up_ptr[i + 1] - up_ptr[i]
            if UNIT:
                if last and router_level:
                    q_state[i] = qf_cur[i] if deg > 0 else ql[i]
                else:
                    q_state[i] = qcur[i]
                    q_full[i] = qf_cur[i]
            else:
                q_state[i] = qcur[i]
        done[b] = j + 1
    assert len(seen) == nb * n_tiles
    if perm is not None:
        out, q_state, q_full = out[:, inv], q_state[inv], q_full[inv]
    return out, q_state, (q_full if UNIT else None)
