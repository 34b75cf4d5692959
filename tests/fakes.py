"""
Oracle-backed stand-ins for the device entry points the router classes use (`Plan.route_host`,
`Plan.runoff_route_host`, `Transform`, the UH / weight transforms): same contracts, everything on the host.  They let
the HOST LOGIC of the routers -- config handling, file I/O, time bookkeeping, path selection, state chaining, basin
sharding -- run in the CPU test tier; the GPU suite runs the same scenarios against the real library.
"""
import numpy as np

import river_route_b200 as rr
from river_route_b200 import plan as plan_mod, transforms
from oracle import oracle


class FakeTransform:
    """Same constructor / methods as transforms.Transform; keeps everything on the host."""

    def __init__(self, indptr, indices, data, n_points, area=None, device=-1):
        self.indptr, self.indices, self.data = np.asarray(indptr), np.asarray(indices), np.asarray(data)
        self.n_rivers, self.n_points, self.area = len(indptr) - 1, int(n_points), area
        self.n_ks, self.kernel, self.state = 0, None, None
        assert self.indices.max(initial=-1) < self.n_points

    def set_unit_hydrograph(self, kernel, state=None):
        self.kernel = np.array(kernel, dtype=np.float64)
        self.state = np.zeros_like(self.kernel) if state is None else np.array(state, dtype=np.float64)
        self.n_ks = self.kernel.shape[0]
        return self

    def uh_state(self):
        return self.state.copy()

    def close(self):
        pass


def _finish(plan, full, out, resample):
    """The output tail the library applies on the device: subset columns, mean over `resample` rows, cast."""
    sub = getattr(plan, '_fake_subset', None)
    if sub is not None:
        full = full[:, sub]
    if resample > 1:
        full = full.reshape(full.shape[0] // resample, resample, full.shape[1]).mean(axis=1)
    out[...] = full.astype(out.dtype)


def _route(plan, mode, q_state, lateral, T, substeps, q_full=None):
    a = plan._fake_arrays
    full = np.zeros((T, plan.n))
    if mode == rr.MODE_RAPID:
        oracle.rapid_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], a['c4_dt'], q_state, lateral, full, substeps)
    elif mode == rr.MODE_MUSKINGUM:
        oracle.muskingum_route(a['indptr'], a['indices'], a['lhs_off'], a['c2'], a['c3'], q_state, full, T, substeps)
    else:
        sp = oracle.unit_split(plan.down.astype(np.int64))
        inner, hw, ai, ah = sp['inner_idx'], sp['hw_idx'], sp['a_inner'], sp['a_hw']
        c1i, c2i, c3i = a['c1'][inner], a['c2'][inner], a['c3'][inner]
        q_ch = q_state[inner].copy()
        q_fu = q_ch.copy() if q_full is None else q_full[inner].copy()
        oracle.unit_route(ai[0], ai[1], -c1i[ai[1]], ai[0], ai[1], ai[2], ah[0], ah[1], ah[2], c1i, c2i, c3i, hw, inner,
                          q_ch, q_fu, lateral, full, substeps)
        if q_full is None:               # router-level call: recombined state (UnitMuskingum.py:94-98)
            q_state[hw] = lateral[-1][hw]
            q_state[inner] = q_fu
        else:                            # kernel-level call: the q_ch / q_full pair of unit_route, headwater entries untouched
            q_state[inner] = q_ch
            q_full[inner] = q_fu
    return full


def install(setattr_fn):
    """``setattr_fn(target, name, value)`` -- pytest's monkeypatch.setattr, or plain setattr in spawned workers."""
    def set_coefficients(self, c1, c2, c3, c4_dt=None):
        indptr, indices = oracle.csc_from_down(self.down)
        self._fake_arrays = dict(indptr=indptr, indices=indices, c1=np.array(c1), c2=np.array(c2), c3=np.array(c3),
                                 c4_dt=None if c4_dt is None else np.array(c4_dt), lhs_off=oracle.lhs_off_data(np.array(c1), indices))

    def set_output_subset(self, indices=None):
        idx = None if indices is None or len(indices) == 0 else np.asarray(indices, dtype=np.int64)
        assert idx is None or (idx.min() >= 0 and idx.max() < self.n)
        self._fake_subset = idx
        self.n_out = self.n if idx is None else int(idx.shape[0])

    def route_host(self, mode, q_state, lateral, out, substeps, q_full=None, resample=1):
        assert out.shape[1] == self.n_out and out.dtype in (np.float32, np.float64) and q_state.shape == (self.n,)
        T = out.shape[0] * resample
        assert mode == rr.MODE_MUSKINGUM or lateral.shape == (T, self.n)
        _finish(self, _route(self, mode, q_state, None if lateral is None else np.ascontiguousarray(lateral, dtype=np.float64),
                             T, substeps, q_full), out, resample)

    def runoff_route_host(self, transform, mode, q_state, runoff, out, substeps, cumulative=False, force_positive=False,
                          as_volumes=False, resample=1):
        T = out.shape[0] * resample
        assert runoff.shape == (T, transform.n_points) and out.shape[1] == self.n_out
        unit = mode == rr.MODE_UNIT
        ql = oracle.weights_transform(transform.indptr, transform.indices, transform.data, runoff, cumulative=cumulative,
                                      force_positive=force_positive, area=transform.area if (as_volumes and not unit) else None)
        if unit:
            ql = oracle.uh_convolve(ql, transform.kernel, transform.state)
        _finish(self, _route(self, mode, q_state, ql, T, substeps), out, resample)

    from river_route_b200 import _lib
    setattr_fn(_lib, 'pinned_empty', lambda shape, dtype=np.float64: np.empty(shape, dtype=dtype))   # no CUDA on this tier
    setattr_fn(plan_mod.Plan, 'set_coefficients', set_coefficients)
    setattr_fn(plan_mod.Plan, 'set_output_subset', set_output_subset)
    def route_ensemble_host(self, mode, q_init, laterals, outs, substeps, resample=1, q_final=None):
        M = len(outs)
        T = outs[0].shape[0] * resample
        finals = np.empty((M, self.n))
        for m in range(M):
            q = (q_init if q_init.ndim == 1 else q_init[m]).copy()
            lat = None if mode == rr.MODE_MUSKINGUM else np.ascontiguousarray(laterals[m], dtype=np.float64)
            _finish(self, _route(self, mode, q, lat, T, substeps), outs[m], resample)
            finals[m] = q
        if q_final is not None:
            q_final[...] = finals
        return np.array(list(finals)).mean(axis=0)

    setattr_fn(plan_mod.Plan, 'route_host', route_host)
    setattr_fn(plan_mod.Plan, 'route_ensemble_host', route_ensemble_host)
    setattr_fn(plan_mod.Plan, 'runoff_route_host', runoff_route_host)
    setattr_fn(transforms, 'Transform', FakeTransform)
    setattr_fn(transforms, 'uh_convolve', lambda lat, ker, st: oracle.uh_convolve(np.ascontiguousarray(lat), ker, st))
    import river_route_b200.uhkernels as uhk
    import river_route_b200.runoff as runoff_mod
    setattr_fn(uhk, 'uh_convolve', lambda lat, ker, st: oracle.uh_convolve(np.ascontiguousarray(lat), ker, st))
    setattr_fn(transforms, 'weights_transform',
                        lambda indptr, indices, data, raw, cumulative=False, force_positive=False, area=None, keep_nan=False:
                        oracle.weights_transform(indptr, indices, data, raw, cumulative=cumulative, force_positive=force_positive,
                                                 area=area, keep_nan=keep_nan))
    setattr_fn(runoff_mod, 'weights_transform', transforms.weights_transform)
