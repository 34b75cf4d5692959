"""Shared test helpers: reference-style array derivation (via the oracle's exact restatements) and tolerances."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402


def network_arrays(down, k, x, dt_routing, dt_runoff=None):
    """CSC + coefficient arrays exactly as the reference derives them (Muskingum.py:172-193)."""
    indptr, indices = oracle.csc_from_down(down)
    c1, c2, c3 = oracle.muskingum_coefficients(k, x, dt_routing)
    d = dict(indptr=indptr, indices=indices, c1=c1, c2=c2, c3=c3, lhs_off=oracle.lhs_off_data(c1, indices))
    if dt_runoff is not None:
        d['c4_dt'] = (c1 + c2) / dt_runoff  # TransformMuskingum.py:104, RapidMuskingum.py:25
    return d


def parity_error(got, ref, col_scale=None):
    """
    SURVEY.md 8d parity measure: |got - ref| <= tol*|ref| + tol*max_t|ref_reach|.
    Returns the smallest tol that makes every element pass.  ``col_scale`` (per reach) adds to the
    per-reach magnitude; used where the reference itself carries noise relative to a larger quantity
    (scipy's FFT convolution leaves ~1e-16 x column-scale residue in entries that are exactly zero).
    """
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref) + (np.max(np.abs(ref), axis=0, keepdims=True) if ref.ndim == 2 else np.max(np.abs(ref)))
    if col_scale is not None:
        scale = scale + np.asarray(col_scale, dtype=np.float64)
    err = np.abs(got - ref)
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.where(scale > 0, err / scale, np.where(err > 0, np.inf, 0.0))
    return float(np.max(ratio)) if ratio.size else 0.0
