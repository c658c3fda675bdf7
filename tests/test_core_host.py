"""The product's C++ analysis layer (csrc/core/stencil.hpp through the C ABI) against
(a) what the real reference generator emitted (tests/golden/ref_emitted.json) and
(b) the oracle's independent restatement.  CPU only -- no compute calls."""
import os
import tempfile

import numpy as np
import pytest

from helpers import SHIPPED, stc_path


def _cases(golden):
    for key, rec in golden.items():
        name, step, dist, extra = key.split("|")
        mf = int(extra.split()[1]) if "--merge-forward" in extra else 5
        yield key, name, int(step.split("=")[1]), int(dist.split("=")[1]), mf, rec


def test_terms_match_reference_emission(built, golden):
    import drstencil_b200 as drs
    for key, name, step, dist, mf, rec in _cases(golden):
        if rec["rc"] != 0:
            continue
        st = drs.Stencil.from_file(stc_path(name)).compose(step)
        terms = st.terms()
        assert [list(t[:3]) for t in terms] == [t[:3] for t in rec["gold_terms"]], key
        assert st.term_texts() == [t[3] for t in rec["gold_terms"]], key
        # value handed to the kernels == value of the printed literal
        assert [t[3] for t in terms] == [float(t[3]) for t in rec["gold_terms"]], key


def test_analysis_matches_reference_macros_and_exit_codes(built, golden):
    import drstencil_b200 as drs
    for key, name, step, dist, mf, rec in _cases(golden):
        st = drs.Stencil.from_file(stc_path(name)).compose(step)
        if rec["rc"] == 1:
            with pytest.raises(drs.DrsError) as e:
                st.analyze(dist, mf)
            assert e.value.code == drs.E_NOREUSE
            assert "No data to reuse" in str(e.value)
            continue
        a = st.analyze(dist, mf)
        m = rec["macros"]
        assert (a["halo"], a["dist"], a["range"]) == (m["Halo"], m["Dist"], m["Range"]), key


def test_partition_sizes_match_oracle(built):
    import drstencil_b200 as drs
    from oracle import oracle
    for name in SHIPPED:
        for step in (1, 2):
            for dist in (0, 1, 2):
                for mf in (1, 5):
                    is3d = name.startswith("3d")
                    s = oracle.parse_stc(stc_path(name), is3d)
                    pts = oracle.compose(s.points, step)
                    _, d = oracle.order_dist(pts, s.dim, dist)
                    part = oracle.partition(pts, s.dim, d, mf)
                    st = drs.Stencil.from_file(stc_path(name)).compose(step)
                    if part is None:
                        with pytest.raises(drs.DrsError):
                            st.analyze(dist, mf)
                        continue
                    a = st.analyze(dist, mf)
                    assert a["forward_slow"] == len(part["forward_slow"])
                    assert a["forward_mid"] == len(part["forward_mid"])
                    assert a["forward_fast"] == len(part["forward_fast"])
                    assert a["backward"] == len(part["backward"])


def test_appendix_c_partition_table(built):
    """SURVEY appendix C (probed from the reference): fwd_j / backward / fwd_i sizes."""
    import drstencil_b200 as drs
    table = {("2d5pt_star", 1): (2, 3, 0), ("2d5pt_star", 2): (5, 8, 0), ("2d9pt_box", 1): (6, 3, 0),
             ("2d9pt_box", 2): (15, 4, 6), ("2d9pt_box", 4): (45, 16, 20), ("2d25pt_box", 1): (15, 4, 6),
             ("2d25pt_box", 2): (45, 16, 20), ("2d9pt_star", 2): (9, 17, 7)}
    for (name, step), (fj, bw, fi) in table.items():
        a = drs.Stencil.from_file(stc_path(name)).compose(step).analyze()
        assert (a["forward_slow"], a["backward"], a["forward_fast"]) == (fj, bw, fi), (name, step)
    a = drs.Stencil.from_file(stc_path("3d7pt_star")).compose(2).analyze()
    assert (a["forward_slow"], a["backward"], a["forward_mid"]) == (7, 13, 5)


def test_parser_quirks(built):
    import drstencil_b200 as drs
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "q.stc")
        # unknown tokens skipped, duplicate point keeps the last coefficient, 2D read ignores L
        open(p, "w").write("L 7\nM 30 junk 5 N 40\niterations 6\nstencil\n0 0 0.5\n1 0 0.25\n0 0 0.125\n-1 0 1e-3\n")
        st = drs.Stencil.from_file(p, is3d=False)
        assert st.shape == (30, 40) and st.iterations == 6
        assert st.terms() == [(0, -1, 0, 0.001), (0, 0, 0, 0.125), (0, 1, 0, 0.25)]
        # coefficients are cut to 6 significant digits on their way to the kernel
        open(p, "w").write("M 8 N 8 iterations 2 stencil 0 0 0.123456789 1 0 0.3333333333 -1 0 1")
        st = drs.Stencil.from_file(p, is3d=False)
        assert [t[3] for t in st.terms()] == [1.0, 0.123457, 0.333333]
        assert st.term_texts() == ["1", "0.123457", "0.333333"]
        with pytest.raises(drs.DrsError) as e:
            drs.Stencil.from_file(os.path.join(td, "missing.stc"), is3d=False)
        assert e.value.code == drs.E_IO and "Error opening stencil file." in str(e.value)


def test_misspelt_iterations(built):
    import drstencil_b200 as drs
    st = drs.Stencil.from_file(stc_path("2d9pt_cross"))
    assert st.iterations == 0 and len(st.terms()) == 9


def test_from_points_equals_from_file(built):
    import drstencil_b200 as drs
    f = drs.Stencil.from_file(stc_path("3d7pt_star"))
    t = f.terms()
    g = drs.Stencil.from_points([x[:3] for x in t], [x[3] for x in t], (16, 16, 16), 4)
    assert g.terms() == t and g.shape == (16, 16, 16)


def test_plan_specialisation_text(built):
    """The generated translation unit carries the gold-order chain with literal coefficients."""
    import drstencil_b200 as drs
    st = drs.Stencil.from_file(stc_path("2d5pt_star")).set_size((64, 64))
    plan = drs.Plan(st, drs.Knobs())
    src = plan.source
    assert "MUL(0, 0, -1, 0.2)" in src and "FMA(0, -1, 0, 0.2)" in src and "FMA(0, 1, 0, 0.2)" in src
    assert src.index("MUL(0, 0, -1") < src.index("FMA(0, -1, 0") < src.index("FMA(0, 0, 0, 0.3)")
    assert "drs_sweep2d.cuh" in src
    info = plan.info
    assert info.halo == 1 and info.kernel_name == "dr_2d5pt_star" and info.tile_x == 128
    # temporal depth 4: base chain in the sweep, composed 81-term chain in the gold kernel
    st = drs.Stencil.from_file(stc_path("2d9pt_box")).set_size((256, 256))
    plan = drs.Plan(st, drs.Knobs(step=4))
    src = plan.source
    assert "#define DRS_TS 4" in src and src.count("FMA(") == 8 + 80
    assert plan.info.halo == 4
    plan = drs.Plan(st, drs.Knobs(step=4, fuse="algebraic"))
    assert "#define DRS_TS 1" in plan.source and plan.source.count("FMA(") == 160


def test_unsupported_descriptions_are_rejected(built):
    import drstencil_b200 as drs
    st = drs.Stencil.from_points([(0, 0), (0, 2), (1, 0), (-1, 0)], [0.1, 0.2, 0.3, 0.4], (32, 32), 2)
    with pytest.raises(drs.DrsError) as e:
        drs.Plan(st, drs.Knobs())
    assert e.value.code == drs.E_ARG


def test_no_gpu_means_loud_failure(built):
    """No CPU fallback: compute entry points fail with DRS_E_NOGPU where there is no device."""
    import ctypes
    import drstencil_b200 as drs
    if drs.lib().drs_device_count() > 0:
        pytest.skip("a GPU is present")
    st = drs.Stencil.from_file(stc_path("2d5pt_star")).set_size((64, 64))
    plan = drs.Plan(st, drs.Knobs())
    buf = ctypes.create_string_buffer(64 * 64 * 8)
    with pytest.raises(drs.DrsError) as e:
        plan.sweep(ctypes.addressof(buf), ctypes.addressof(buf) + 8)
    assert e.value.code == drs.E_NOGPU


def test_data_reuse_mode_plans_and_refusals(built):
    """`--fuse reuse`: the plan carries the reference's partition as three ordered partial sums, and
    refuses what the reference refuses, with its messages and the matching error codes."""
    import pytest
    import drstencil_b200 as drs
    from helpers import stc_path

    def plan(name, shape, **kn):
        return drs.Plan(drs.Stencil.from_file(stc_path(name)).set_size(shape), drs.Knobs(fuse="reuse", **kn))

    p = plan("2d5pt_star", (64, 64))
    src = p.source
    assert "drs_reuse.cuh" in src and "#define DRS_DIST 1" in src and "#define DRS_HAS_FWD_FAST 0" in src
    # appendix B of SURVEY.md: out[j+1] = 0.2*in[j] + 0.3*in[j+1] (operands relative to the sweep centre j);
    # backward = 0.2*in[j][i-1] + 0.2*in[j][i+1] + 0.2*in[j+1][i]
    fwd = src[src.index("#define DRS_FWD_SLOW"):src.index("#define DRS_HAS_BWD")]
    assert "MUL(0, 1, 0, 0.3)" in fwd and "FMA(0, 0, 0, 0.2)" in fwd and fwd.count("FMA(0") == 1
    bwd = src[src.index("#define DRS_BWD"):src.index("#define DRS_HAS_FWD_MID")]
    assert bwd.count("MUL(0") == 1 and bwd.count("FMA(0") == 2
    p = plan("2d25pt_box", (64, 64))
    assert "#define DRS_DIST 2" in p.source and "#define DRS_HAS_FWD_FAST 1" in p.source          # 15 / 4 / 6 (appendix C)
    fast = p.source[p.source.index("#define DRS_FWD_FAST("):].split("#define")[1]
    assert fast.count("MUL(0") + fast.count("FMA(0") == 6
    p = plan("3d7pt_star", (32, 32, 32), step=2, merge_forward=1)
    assert "#define DRS_HAS_FWD_MID 1" in p.source
    with pytest.raises(drs.DrsError, match="No data to reuse") as e:
        plan("2d5pt_cross", (64, 64))                       # drstencil_2d.hpp:217-220, exit 1
    assert e.value.code == drs.E_NOREUSE
    with pytest.raises(drs.DrsError, match="Invalid configuration") as e:
        plan("2d25pt_box", (64, 64), bx=4)                  # codegen_2d.hpp:52-56
    assert e.value.code == drs.E_CONFIG
