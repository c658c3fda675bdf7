"""Pins the oracle's host-side restatement (oracle/oracle.py) to the REAL reference generator:
tests/golden/ref_emitted.json holds what /root/reference/main.cpp emitted for every shipped
stencil x step x dist (made by tests/golden/make_golden.py).  CPU only."""
import pytest

from helpers import SHIPPED, stc_path


def _cases(golden):
    for key, rec in golden.items():
        name, step, dist, extra = key.split("|")
        step = int(step.split("=")[1])
        dist = int(dist.split("=")[1])
        mf = 5
        if "--merge-forward" in extra:
            mf = int(extra.split()[1])
        yield key, name, step, dist, mf, rec


def test_fixture_covers_every_shipped_stencil(golden):
    names = {k.split("|")[0] for k in golden}
    assert names == set(SHIPPED)
    assert len(golden) >= 60


def test_gold_expression_order_offsets_and_literals(golden):
    from oracle import oracle
    n = 0
    for key, name, step, dist, mf, rec in _cases(golden):
        if rec["rc"] != 0:
            continue
        is3d = name.startswith("3d")
        s = oracle.parse_stc(stc_path(name), is3d)
        pts = oracle.compose(s.points, step)
        keys = sorted(pts)
        assert [list(k) for k in keys] == [t[:3] for t in rec["gold_terms"]], key
        assert [oracle.literal_text(pts[k]) for k in keys] == [t[3] for t in rec["gold_terms"]], key
        n += 1
    assert n >= 40


def test_literal_roundtrip_changes_bits_for_deep_fusion(golden):
    """most coefficients of 2d9pt_box step 4 differ from their printed literal."""
    from oracle import oracle
    s = oracle.parse_stc(stc_path("2d9pt_box"), False)
    pts = oracle.compose(s.points, 4)
    assert len(pts) == 81
    changed = sum(1 for c in pts.values() if oracle.literal(c) != c)
    assert changed >= 60


def test_macros_and_exit_codes(golden):
    from oracle import oracle
    for key, name, step, dist, mf, rec in _cases(golden):
        is3d = name.startswith("3d")
        s = oracle.parse_stc(stc_path(name), is3d)
        pts = oracle.compose(s.points, step)
        halo, d = oracle.order_dist(pts, s.dim, dist)
        part = oracle.partition(pts, s.dim, d, mf)
        if rec["rc"] == 1:
            assert part is None, key
            assert "No data to reuse" in rec["stdout"]
            continue
        assert rec["rc"] == 0 and part is not None, key
        m = rec["macros"]
        assert m["Halo"] == halo and m["Dist"] == d, key
        assert m["Range"] == part["high"] - part["low"] + 1, key
        assert m["M"] == s.M and m["N"] == s.N, key
        if is3d:
            assert m["L"] == s.L
        if name != "2d9pt_cross":   # misspelt key: the reference emits an uninitialised int there
            assert m["Iterations"] == s.iterations, key


def test_misspelt_iterations_key_is_ignored():
    from oracle import oracle
    s = oracle.parse_stc(stc_path("2d9pt_cross"), False)
    assert s.iterations == 0 and len(s.points) == 9


def test_sweep_count_matches_emitted_loop():
    from oracle import oracle
    assert oracle.sweep_count(4, 1) == 4
    assert oracle.sweep_count(4, 2) == 2
    assert oracle.sweep_count(10, 1) == 10
    assert oracle.sweep_count(10, 2) == 6    # t = 0, 4, 8
    assert oracle.sweep_count(8, 4) == 2
    assert oracle.sweep_count(0, 1) == 0
