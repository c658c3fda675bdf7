"""Host logic of the streamed drs_run_host (csrc/core/host_schedule.hpp), checked WITHOUT a GPU:
the step list the C ABI reports (drs_plan_host_schedule) is replayed with the CPU oracle standing in
for the sweep kernel.  Device planes that were not uploaded yet are poisoned with NaN, so a sweep
that runs too early, a download that comes too late (after a later sweep overwrote final planes) or
an upload that reads host planes a download already replaced all show up as a mismatch against the
plain schedule (`for t: sweep(A,B); sweep(B,A)`, /root/reference/codegen_2d.hpp:610-613)."""
import numpy as np
import pytest

from helpers import oracle_terms, stc_path

UPLOAD, SWEEP, DOWNLOAD = 0, 1, 2


def _plan(name, shape, block, **kn):
    import drstencil_b200 as drs
    st = drs.Stencil.from_file(stc_path(name)).set_size(shape)
    plan = drs.Plan(st, drs.Knobs(**kn))
    plan.set_host_block(block)
    return plan


def _replay(steps, order, host, offs, coefs, halo):
    """Executes the step list in `order` on emulated device buffers; `host` is updated in place."""
    from oracle import oracle
    dev = [np.full(host.shape, np.nan), np.zeros(host.shape)]     # A: nothing uploaded yet; B: cleared
    for i in order:
        kind, block, sweep, lo, hi = steps[i]
        if kind == UPLOAD:
            dev[0][lo:hi] = host[lo:hi]
        elif kind == DOWNLOAD:
            host[lo:hi] = dev[0][lo:hi]
        else:
            src, dst = (dev[0], dev[1]) if sweep & 1 else (dev[1], dev[0])
            full = dst.copy()
            oracle.sweep(src, full, offs, coefs, halo)            # whole interior ...
            dst[lo:hi] = full[lo:hi]                              # ... of which the launch writes [lo, hi)
    return dev


def _orders(steps):
    """The two extreme interleavings the stream/event chain allows:
    lazy uploads + eager downloads (list order), eager uploads + downloads deferred to the end."""
    idx = list(range(len(steps)))
    ups = [i for i in idx if steps[i][0] == UPLOAD]
    dns = [i for i in idx if steps[i][0] == DOWNLOAD]
    mid = [i for i in idx if steps[i][0] == SWEEP]
    return {"lazy-up/eager-down": idx, "eager-up/late-down": ups + mid + dns}


CASES = [
    # name, shape, knobs, iterations, block
    ("2d5pt_star", (60, 24), dict(), 10, 2),
    ("2d5pt_star", (61, 24), dict(), 10, 7),
    ("2d5pt_star", (40, 24), dict(), 30, 6),            # n * Halo > block thickness
    ("2d9pt_star", (70, 20), dict(), 6, 4),             # radius 2
    ("2d9pt_box", (90, 24), dict(step=2, fuse="algebraic"), 8, 9),
    ("2d9pt_box", (120, 24), dict(step=4, fuse="algebraic"), 16, 8),
    ("3d7pt_star", (40, 10, 12), dict(), 8, 5),
    ("3d7pt_star", (33, 10, 12), dict(step=2, fuse="algebraic"), 8, 4),
    ("3d9pt_cross", (26, 10, 12), dict(), 4, 3),
    # a remainder thinner than 2*Halo joins its predecessor (round 1: slow = S + 1 with S = 2*Halo left a block of one
    # row, below the thickness the schedule's correctness argument needs)
    ("2d9pt_star", (9, 20), dict(), 6, 4),              # Halo 2, S = 2*Halo = 4, slow = 2*S + 1
    ("2d5pt_star", (5, 24), dict(), 6, 2),              # Halo 1, S = 2, slow = 2*S + 1
    ("3d7pt_star", (7, 10, 12), dict(step=2, fuse="algebraic"), 8, 4),   # Halo 2, S = 4, slow = S + 3
]


@pytest.mark.parametrize("name,shape,kn,iters,block", CASES)
def test_step_list_is_a_legal_order_of_the_plain_schedule(built, name, shape, kn, iters, block):
    from oracle import oracle
    step = kn.get("step", 1)
    plan = _plan(name, shape, block, **kn)
    steps = plan.host_schedule(iters)
    assert steps, "expected the streamed path for this case"
    offs, coefs, halo = oracle_terms(name, step)
    assert halo == plan.halo
    n = oracle.sweep_count(iters, step)
    slow = shape[0]
    # structure: blocks tile the slow axis, every block >= 2*Halo thick, sweeps partition the interior
    ups = [s for s in steps if s[0] == UPLOAD]
    assert ups[0][3] == 0 and ups[-1][4] == slow
    assert all(a[4] == b[3] for a, b in zip(ups, ups[1:]))
    assert min(u[4] - u[3] for u in ups) >= 2 * halo
    for s in range(1, n + 1):
        rng = sorted((lo, hi) for k, b, sw, lo, hi in steps if k == SWEEP and sw == s)
        assert rng[0][0] == halo and rng[-1][1] == slow - halo
        assert all(a[1] == b[0] for a, b in zip(rng, rng[1:]))
    dn = [(lo, hi) for k, b, sw, lo, hi in steps if k == DOWNLOAD and hi > lo]
    assert dn[0][0] == 0 and dn[-1][1] == slow and all(a[1] == b[0] for a, b in zip(dn, dn[1:]))
    # semantics
    a0 = oracle.lcg_array(shape, np.float64, 5)
    refA, refB = a0.copy(), np.zeros(shape)
    assert oracle.run(refA, refB, offs, coefs, halo, iters, step) == n
    for label, order in _orders(steps).items():
        host = a0.copy()
        dev = _replay(steps, order, host, offs, coefs, halo)
        assert np.array_equal(host, refA), (label, name, shape, block)
        assert np.array_equal(dev[0], refA) and np.array_equal(dev[1], refB), label


def test_plain_sequence_when_streaming_does_not_apply(built):
    """One block would cover the grid, overlap switched off, slab plans: no step list."""
    assert _plan("2d5pt_star", (60, 24), 60).host_schedule(10) == []
    assert _plan("2d5pt_star", (60, 24), -1).host_schedule(10) == []
    assert _plan("2d5pt_star", (60, 24), 4).host_schedule(0) == []
    auto = _plan("2d5pt_star", (60, 24), 0)          # 60 rows of 192 bytes: far below 32 MiB per block
    assert auto.host_schedule(10) == []
    big = _plan("3d7pt_star", (1536, 1536, 1536), 0)  # c5: engine-chosen blocks, nothing allocated
    steps = big.host_schedule(100)
    ups = [s for s in steps if s[0] == UPLOAD]
    assert 8 <= len(ups) <= 20 and ups[0][4] - ups[0][3] < ups[1][4] - ups[1][3]
    assert sum(1 for s in steps if s[0] == SWEEP) <= 100 * len(ups)
