"""TEST INFRASTRUCTURE ONLY -- Python side of the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; nothing under drstencil_b200/ does.

An independent (pure Python + numpy) restatement of the reference's host-side analysis, used to
check the product's C++ core, plus a ctypes wrapper over oracle/libdrs_oracle.so for the sweeps:

  parse_stc        /root/reference/drstencil_2d.hpp:48-73, drstencil.hpp:52-78
  compose          /root/reference/drstencil_2d.hpp:231-251, drstencil.hpp:262-282
  literal          /root/reference/drstencil_2d.hpp:174 (default ostream formatting == "%g")
  order_dist       /root/reference/drstencil_2d.hpp:82-97, drstencil.hpp:88-103
  partition        /root/reference/drstencil_2d.hpp:180-228, drstencil.hpp:198-259
  sweeps/schedule  see oracle/drs_oracle.c

Parity status: pinned -- see the header of oracle/drs_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/drs_oracle.c -> oracle/libdrs_oracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "libdrs_oracle.so")
    src = os.path.join(_HERE, "drs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
             "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libdrs_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        ll, i32, vp = ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p
        L.drs_oracle_threads.restype = i32
        L.drs_oracle_set_threads.argtypes = [i32]
        L.drs_oracle_set_threads.restype = None
        L.drs_oracle_fill_rand.argtypes = [vp, ctypes.c_size_t, i32]
        L.drs_oracle_fill_rand_f32.argtypes = [vp, ctypes.c_size_t, i32]
        L.drs_oracle_fill_lcg.argtypes = [vp, ctypes.c_size_t, ctypes.c_uint]
        L.drs_oracle_fill_lcg_f32.argtypes = [vp, ctypes.c_size_t, ctypes.c_uint]
        for name in ("drs_oracle_sweep_f64", "drs_oracle_sweep_f32"):
            getattr(L, name).argtypes = [i32, ll, ll, ll, i32, i32, vp, vp, vp, vp, i32]
            getattr(L, name).restype = None
        for name in ("drs_oracle_run_f64", "drs_oracle_run_f32"):
            getattr(L, name).argtypes = [i32, ll, ll, ll, i32, i32, vp, vp, vp, vp, i32, i32, i32]
            getattr(L, name).restype = i32
        L.drs_oracle_check_error.argtypes = [i32, ll, ll, ll, i32, vp, vp, vp]
        _LIB = L
    return _LIB


def set_threads(n: int) -> int:
    """OpenMP threads of the sweeps from now on (bench.py: all host cores even under torchrun's OMP_NUM_THREADS=1)."""
    lib().drs_oracle_set_threads(int(n))
    return lib().drs_oracle_threads()


# --------------------------------------------------------------------------------------------
# host-side analysis, restated
# --------------------------------------------------------------------------------------------

@dataclass
class Stc:
    dim: int
    L: int = 1
    M: int = 0
    N: int = 0
    iterations: int = 0
    points: dict = field(default_factory=dict)  # {(k, j, i): coef}, k == 0 in 2D


def _is_int(tok: str) -> bool:
    try:
        int(tok)
        return True
    except ValueError:
        return False


def _is_float(tok: str) -> bool:
    try:
        float(tok)
        return True
    except ValueError:
        return False


def parse_stc(path: str, is3d: bool) -> Stc:
    """Whitespace token stream; keys L(3D) M N iterations stencil; unknown tokens skipped;
    tuples after `stencil` until the first non-numeric token; repeated point -> last wins."""
    with open(path) as f:
        toks = f.read().split()
    s = Stc(dim=3 if is3d else 2)
    n = 0
    while n < len(toks):
        t = toks[n]
        n += 1
        if t in ("M", "N", "iterations") or (is3d and t == "L"):
            if n < len(toks) and _is_int(toks[n]):
                setattr(s, t, int(toks[n]))
                n += 1
        elif t == "stencil":
            w = 4 if is3d else 3
            while n + w <= len(toks) and all(_is_int(x) for x in toks[n:n + w - 1]) and _is_float(toks[n + w - 1]):
                ints = [int(x) for x in toks[n:n + w - 1]]
                key = tuple(ints) if is3d else (0, ints[0], ints[1])
                s.points[key] = float(toks[n + w - 1])
                n += w
            break
    return s


def compose(points: dict, step: int) -> dict:
    """Composed operator; depth-first over sorted base points, product left to right from 1.0,
    accumulation in visit order (this fixes the floating-point value of every coefficient)."""
    base = sorted(points.items())
    acc: dict = {}

    def walk(at, prod, left):
        if left == 0:
            if at in acc:
                acc[at] += prod
            else:
                acc[at] = prod
            return
        for (k, j, i), c in base:
            walk((at[0] + k, at[1] + j, at[2] + i), prod * c, left - 1)

    walk((0, 0, 0), 1.0, step)
    return acc


def literal_text(c: float) -> str:
    return "%g" % c


def literal(c: float) -> float:
    return float(literal_text(c))


def terms(points: dict):
    """(offsets int32[P,3], coefs float64[P]) in evaluation (ascending key) order, coefficients
    as the values of their printed literals."""
    keys = sorted(points)
    offs = np.array(keys, dtype=np.int32).reshape(-1, 3)
    coefs = np.array([literal(points[k]) for k in keys], dtype=np.float64)
    return offs, coefs


def order_dist(points: dict, dim: int, dist_opt: int = 0):
    ax = 0 if dim == 3 else 1
    hi = max([0] + [p[ax] for p in points])
    lo = min([0] + [p[ax] for p in points])
    return hi, (dist_opt if dist_opt != 0 else (hi - lo) >> 1)


def partition(points: dict, dim: int, dist: int, merge_forward: int = 5):
    """Returns dict(forward_slow, forward_mid, forward_fast, backward, low, high) or None when the
    slow-axis forward set is empty (reference: "No data to reuse", exit 1)."""
    ax_slow = 0 if dim == 3 else 1

    def sh(p, axis, by):
        q = list(p)
        q[axis] += by
        return tuple(q)

    keys = sorted(points)
    done = set()
    fwd_slow, fwd_mid, fwd_fast, back = set(), set(), set(), set()
    for p in keys:
        q = sh(p, ax_slow, -dist)
        if q in points:
            fwd_slow.add(p)
            done.add(q)
    if dim == 3:
        for p in keys:
            q = sh(p, 1, -dist)
            if q in points and q not in done:
                fwd_mid.add(p)
                done.add(q)
    for p in keys:
        q = sh(p, 2, -dist)
        if q in points and q not in done:
            fwd_fast.add(p)
            done.add(q)
    for p in keys:
        if p not in done:
            back.add(p)
            done.add(p)
    if not fwd_slow:
        return None
    if dim == 3 and len(fwd_mid) < merge_forward:
        back |= {sh(p, 1, -dist) for p in fwd_mid}
        fwd_mid = set()
    if len(fwd_fast) < merge_forward:
        back |= {sh(p, 2, -dist) for p in fwd_fast}
        fwd_fast = set()
    allp = fwd_slow | fwd_mid | fwd_fast | back
    low = min([1] + [p[ax_slow] for p in allp])
    high = max([-1] + [p[ax_slow] for p in allp])
    return dict(forward_slow=fwd_slow, forward_mid=fwd_mid, forward_fast=fwd_fast, backward=back,
                low=low, high=high)


def sweep_count(iterations: int, step: int) -> int:
    """for (t = 0; t < Iterations; t += 2*step) { A->B; B->A; }"""
    n = 0
    t = 0
    while t < iterations:
        n += 2
        t += 2 * step
    return n


# --------------------------------------------------------------------------------------------
# sweeps
# --------------------------------------------------------------------------------------------

def _shape3(a: np.ndarray):
    if a.ndim == 2:
        return 2, 1, a.shape[0], a.shape[1]
    return 3, a.shape[0], a.shape[1], a.shape[2]


def rand_array(shape, dtype=np.float64, reseed=True) -> np.ndarray:
    """The reference's input: unseeded glibc rand()/(RAND_MAX-1), row-major."""
    a = np.empty(shape, dtype=dtype)
    fn = lib().drs_oracle_fill_rand if dtype == np.float64 else lib().drs_oracle_fill_rand_f32
    fn(a.ctypes.data, a.size, 1 if reseed else 0)
    return a


def lcg_array(shape, dtype=np.float64, seed=0) -> np.ndarray:
    a = np.empty(shape, dtype=dtype)
    fn = lib().drs_oracle_fill_lcg if dtype == np.float64 else lib().drs_oracle_fill_lcg_f32
    fn(a.ctypes.data, a.size, seed)
    return a


def sweep(inp: np.ndarray, out: np.ndarray, offs: np.ndarray, coefs: np.ndarray, halo: int,
          contract: bool = True) -> None:
    """out[interior] = gold expression of inp; the halo ring of `out` keeps its old contents."""
    assert inp.flags.c_contiguous and out.flags.c_contiguous and inp.shape == out.shape
    assert inp.dtype == out.dtype and inp.dtype in (np.float64, np.float32)
    dim, L, M, N = _shape3(inp)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    fn = lib().drs_oracle_sweep_f64 if inp.dtype == np.float64 else lib().drs_oracle_sweep_f32
    fn(dim, L, M, N, halo, len(coefs), offs.ctypes.data, coefs.ctypes.data, inp.ctypes.data,
       out.ctypes.data, 1 if contract else 0)


def sweep_reuse(inp: np.ndarray, out: np.ndarray, points: dict, dim: int, dist_opt: int = 0, merge_forward: int = 5) -> bool:
    """The reference's data-reuse evaluation of one sweep (its dr_<name> kernel, codegen_2d.hpp:345-366,
    codegen.hpp:391-427): out = ((forward_slow + backward) + forward_mid) + forward_fast on the interior, every
    partial sum being the contracted chain of its own terms (a forward term reads the input at its point's offset with
    the coefficient of the point `Dist` earlier along its axis: gen_forward_j, drstencil_2d.hpp:120-135).  The
    reference's two cross-thread atomics can land in either order; this is the order the engine's A/B kernel fixes.
    Returns False where the reference exits with "No data to reuse"."""
    halo, dist = order_dist(points, dim, dist_opt)
    part = partition(points, dim, dist, merge_forward)
    if part is None:
        return False
    ax_slow = 0 if dim == 3 else 1

    def partial(pts, axis):
        keys = sorted(pts)
        if not keys:
            return None
        # seen from the OUTPUT point a forward term p is simply the stencil point q = p - Dist*e with its own
        # coefficient (the sweep is centred Dist rows before the output it forwards to); the partition decides
        # the grouping and the order of the sum, not the operands
        qs = []
        for p in keys:
            q = list(p)
            if axis is not None:
                q[axis] -= dist
            qs.append(tuple(q))
        offs = np.array(qs, dtype=np.int32).reshape(-1, 3)
        co = np.array([literal(points[q]) for q in qs], dtype=np.float64)
        tmp = np.zeros_like(inp)
        sweep(inp, tmp, offs, co, halo)
        return tmp

    inner = tuple(slice(halo, n - halo) for n in inp.shape)
    acc = partial(part["forward_slow"], ax_slow)
    for pts, axis in ((part["backward"], None), (part["forward_mid"], 1), (part["forward_fast"], 2)):
        t = partial(pts, axis)
        if t is not None:
            acc = acc + t          # IEEE round-to-nearest add of the rounded partial sums (atomicAdd / RED.ADD)
    out[inner] = acc[inner]
    return True


def run(A: np.ndarray, B: np.ndarray, offs, coefs, halo: int, iterations: int, step: int,
        contract: bool = True) -> int:
    """The emitted program's ping-pong schedule; result ends in A.  Returns sweeps done."""
    dim, L, M, N = _shape3(A)
    offs = np.ascontiguousarray(offs, dtype=np.int32)
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    fn = lib().drs_oracle_run_f64 if A.dtype == np.float64 else lib().drs_oracle_run_f32
    return fn(dim, L, M, N, halo, len(coefs), offs.ctypes.data, coefs.ctypes.data, A.ctypes.data,
              B.ctypes.data, iterations, step, 1 if contract else 0)


def sweep_numpy(inp: np.ndarray, out: np.ndarray, offs, coefs, halo: int) -> None:
    """Slow cross-check of the C sweep for small cases: same chain with exact FMA emulated
    through Python's math.fma when present, else via float128-free splitting (not needed for
    tests that only compare structure).  fp64 only."""
    import math

    dim, L, M, N = _shape3(inp)
    a = inp.reshape(L, M, N)
    o = out.reshape(L, M, N)
    P = len(coefs)
    ks = range(halo, L - halo) if dim == 3 else range(1)
    fma = getattr(math, "fma", None)
    for k in ks:
        for j in range(halo, M - halo):
            for i in range(halo, N - halo):
                v = [float(a[k + int(d[0]), j + int(d[1]), i + int(d[2])]) for d in offs]
                if P == 1:
                    acc = float(coefs[0]) * v[0]
                else:
                    acc = float(coefs[1]) * v[1]
                    order = [0] + list(range(2, P))
                    for q in order:
                        if fma is not None:
                            acc = fma(float(coefs[q]), v[q], acc)
                        else:  # exact product+sum in rational arithmetic, rounded once
                            from fractions import Fraction
                            acc = float(Fraction(float(coefs[q])) * Fraction(v[q]) + Fraction(acc))
                o[k, j, i] = acc


def check_error(out: np.ndarray, ref: np.ndarray, halo: int):
    """(max_abs_error with the reference's 1e-13 floor, rms, (k, j, i))."""
    dim, L, M, N = _shape3(out)
    res = np.zeros(5)
    lib().drs_oracle_check_error(dim, L, M, N, halo, np.ascontiguousarray(out, dtype=np.float64).ctypes.data,
                                 np.ascontiguousarray(ref, dtype=np.float64).ctypes.data, res.ctypes.data)
    return res[0], res[1], (int(res[2]), int(res[3]), int(res[4]))
