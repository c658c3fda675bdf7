"""Shared helpers for the parity tests (test infrastructure: may use oracle/)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STC = os.path.join(ROOT, "stc")

SHIPPED = ["2d5pt_star", "2d5pt_cross", "2d9pt_star", "2d9pt_box", "2d9pt_cross", "2d25pt_box",
           "3d7pt_star", "3d9pt_cross"]


def stc_path(name):
    return os.path.join(STC, name + ".stc")


def oracle_terms(name, step):
    """(offs, coefs, halo) of the composed operator from the ORACLE's own restatement."""
    from oracle import oracle
    is3d = name.startswith("3d")
    s = oracle.parse_stc(stc_path(name), is3d)
    pts = oracle.compose(s.points, step)
    offs, coefs = oracle.terms(pts)
    halo, _ = oracle.order_dist(pts, s.dim)
    return offs, coefs, halo


def oracle_run(name, step, shape, sweeps, dtype=np.float64, a0=None):
    """A after `sweeps` alternating gold sweeps of the composed operator starting from the
    reference's rand() input (or a0) and a zero second buffer.  Returns (A, B)."""
    from oracle import oracle
    offs, coefs, halo = oracle_terms(name, step)
    A = oracle.rand_array(shape, dtype) if a0 is None else a0.copy()
    B = np.zeros(shape, dtype)
    bufs = [A, B]
    for s in range(sweeps):
        oracle.sweep(bufs[s & 1], bufs[(s & 1) ^ 1], offs, coefs, halo)
    return A, B


def max_rel(x, ref):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300))


def pointwise_rel(x, ref, floor_frac=1e-6):
    """max over points of |x - ref| / max(|ref|, floor), floor = floor_frac * max |ref| (fields grow like
    (sum of coefficients)^n and span many decades between the frozen ring and the centre; a global max-abs
    measure alone says nothing about the small values)."""
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    scale = float(np.max(np.abs(ref)))
    if scale == 0.0:
        return float(np.max(np.abs(x)))
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), floor_frac * scale)))
