"""The C sweep of the oracle against an exact (rational-arithmetic) evaluation of the same
mul/fma chain on tiny grids, plus the input generator's known first values.  CPU only."""
import numpy as np
import pytest

from helpers import oracle_terms


def test_rand_input_known_answers(built):
    """first draws of the never-seeded glibc generator over RAND_MAX-1."""
    from oracle import oracle
    a = oracle.rand_array((1, 4))
    assert a[0, 0] == 1804289383 / 2147483646
    assert a[0, 1] == 846930886 / 2147483646
    assert a[0, 2] == 1681692777 / 2147483646
    b = oracle.rand_array((2, 2))
    assert np.array_equal(a.ravel(), b.ravel())     # reseeded -> same sequence


@pytest.mark.parametrize("name,step,shape", [
    ("2d5pt_star", 1, (9, 11)), ("2d9pt_box", 2, (11, 12)), ("2d25pt_box", 1, (9, 10)),
    ("2d9pt_cross", 1, (10, 9)), ("3d7pt_star", 1, (5, 6, 7)), ("3d9pt_cross", 1, (5, 5, 6)),
    ("3d7pt_star", 2, (7, 7, 8)),
])
def test_c_sweep_equals_exact_chain(built, name, step, shape):
    from oracle import oracle
    offs, coefs, halo = oracle_terms(name, step)
    a = oracle.rand_array(shape)
    out_c = np.full(shape, -7.0)
    out_py = np.full(shape, -7.0)
    oracle.sweep(a, out_c, offs, coefs, halo)
    oracle.sweep_numpy(a, out_py, offs, coefs, halo)
    assert np.array_equal(out_c, out_py)
    # the ring is untouched
    ring = np.ones(shape, bool)
    ring[tuple(slice(halo, n - halo) for n in shape)] = False
    assert np.all(out_c[ring] == -7.0)
    assert not np.any(out_c[~ring] == -7.0)


def test_contracted_and_uncontracted_differ_only_in_rounding(built):
    from oracle import oracle
    offs, coefs, halo = oracle_terms("2d9pt_box", 1)
    a = oracle.rand_array((64, 64))
    o1, o2 = np.zeros_like(a), np.zeros_like(a)
    oracle.sweep(a, o1, offs, coefs, halo, contract=True)
    oracle.sweep(a, o2, offs, coefs, halo, contract=False)
    assert np.max(np.abs(o1 - o2)) < 1e-15
    assert not np.array_equal(o1, o2)


def test_f32_sweep_tracks_f64(built):
    from oracle import oracle
    offs, coefs, halo = oracle_terms("2d25pt_box", 1)
    a = oracle.rand_array((40, 48))
    o64 = np.zeros_like(a)
    oracle.sweep(a, o64, offs, coefs, halo)
    a32 = a.astype(np.float32)
    o32 = np.zeros_like(a32)
    oracle.sweep(a32, o32, offs, coefs, halo)
    assert np.max(np.abs(o32 - o64)) / np.max(np.abs(o64)) < 1e-6


def test_run_schedule_result_in_a(built):
    from oracle import oracle
    offs, coefs, halo = oracle_terms("2d5pt_star", 1)
    A = oracle.rand_array((20, 24))
    B = np.zeros_like(A)
    A0 = A.copy()
    n = oracle.run(A, B, offs, coefs, halo, iterations=4, step=1)
    assert n == 4
    x, y = A0.copy(), np.zeros_like(A0)
    for s in range(4):
        src, dst = (x, y) if s % 2 == 0 else (y, x)
        oracle.sweep(src, dst, offs, coefs, halo)
    assert np.array_equal(A, x) and np.array_equal(B, y)


def test_fused_sweep_equals_substeps_in_the_interior(built):
    """one composed sweep == `step` base sub-steps wherever the footprint is in
    bounds (no ring freeze at sub-steps), up to rounding."""
    from oracle import oracle
    offs1, coefs1, h1 = oracle_terms("2d9pt_box", 1)
    offs4, coefs4, h4 = oracle_terms("2d9pt_box", 4)
    assert h4 == 4 * h1
    a = oracle.rand_array((40, 44))
    fused = np.zeros_like(a)
    oracle.sweep(a, fused, offs4, coefs4, h4)
    cur = a.copy()
    for s in range(4):
        nxt = np.zeros_like(cur)
        oracle.sweep(cur, nxt, offs1, coefs1, h1)
        cur = nxt
    inner = (slice(4, -4), slice(4, -4))
    assert np.max(np.abs(cur[inner] - fused[inner])) / np.max(np.abs(fused[inner])) < 1e-14
