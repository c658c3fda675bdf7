"""Tuner rebuild: configuration-name grammar (reference tuning.py:72-86), search-space filter and
the name-keyed Nsight Compute parser, on CSV captured from Nsight Compute 2025.2 on the B200 box
(tests/golden/ncu_*_sample.csv).  CPU only."""
import os

from helpers import ROOT


def test_config_name_grammar_matches_reference():
    from drstencil_b200.tuner.space import Config, cfg_to_command_line, cfg_to_string
    # streaming: fu{step}d{dist}bx{bx}sn{sn}u{unroll}(bmx|cmx){m}mf{t}[p]
    c = Config(step=2, dist=2, bx=128, streaming=True, sn=64, s_unroll=4, block_merge_x=False, mx=2, prefetch=True)
    assert cfg_to_string(c) == "fu2d2bx128sn64u4cmx2mf5p"
    assert cfg_to_command_line(c) == (" --step 2 --dist 2 --bx 128 --streaming --sn 64 --stream-unroll 4"
                                      " --cyclic-merge-y 1 --cyclic-merge-x 2 --merge-forward 5 --prefetch")
    # non-streaming: fu..d..bx..y..(bmx|cmx)m(bmy|cmy)m mf
    c = Config(step=1, dist=1, bx=32, by=8, streaming=False, block_merge_x=True, mx=2, block_merge_y=False, my=4)
    assert cfg_to_string(c) == "fu1d1bx32y8bmx2cmy4mf5"
    # 3D (benchmarks/3d7pt_star/tuning.py:57-72): block shape, sn and unroll always named -- two configurations that
    # differ only in sn must not share a name (result files and --resume are keyed on it)
    c = Config(step=1, bx=32, by=2, streaming=False, sn=64, s_unroll=4, block_merge_y=True, my=1, rows_3d=6, dim=3)
    assert cfg_to_string(c) == "fu1d0bx32y2sn64u4bmx1bmy1mf5ry6"
    assert cfg_to_command_line(c) == (" --step 1 --dist 0 --bx 32 --by 2 --sn 64 --stream-unroll 4 --block-merge-y 1"
                                      " --block-merge-x 1 --merge-forward 5 --rows-3d 6")
    from drstencil_b200.tuner.space import search_space
    for experimental in (False, True):
        names = [cfg_to_string(x) for x in search_space(3, 1, experimental=experimental)]
        assert len(names) == len(set(names))
    shared = [x for x in search_space(3, 1, experimental=True) if x.share_x * x.share_y > 1]
    assert shared and cfg_to_string(shared[0]).endswith("sx%dsy%d" % (shared[0].share_x, shared[0].share_y))
    assert " --share-x " in cfg_to_command_line(shared[0])
    # engine axes are appended only when they differ from the defaults
    c = Config(step=4, bx=64, sn=256, mx=2, stages=8, min_blocks=4, dtype="f32", fuse="algebraic")
    assert cfg_to_string(c) == "fu4d0bx64sn256u4bmx2mf5st8mb4f32alg"


def test_config_names_round_trip_and_presets_are_tuner_names():
    """A result-file name denotes exactly one configuration (cfg_from_string o cfg_to_string = id over the whole
    space), and every BASELINE preset is stated as such a name (presets.TUNED)."""
    from drstencil_b200.tuner.space import cfg_from_string, cfg_to_string, search_space
    for dim, radius, step, dtype in ((2, 1, 1, "f64"), (2, 1, 4, "f64"), (2, 2, 1, "f32"), (3, 1, 1, "f64"), (3, 1, 2, "f64")):
        for c in search_space(dim, radius, step, dtype):
            assert cfg_from_string(cfg_to_string(c), dim) == c
    from drstencil_b200.presets import PRESETS, TUNED
    for wl, (path, dim, name) in TUNED.items():
        c = cfg_from_string(name, dim)
        assert repr(c.knobs()) == repr(PRESETS[wl][1])
    import pytest
    with pytest.raises(ValueError):
        cfg_from_string("fu1d0bx64sn128u4bmx2mf5st2zz", 2)


def test_search_space_filter():
    from drstencil_b200.tuner.space import Config, filter_config, search_space
    sp = search_space(2, 1, step=4)
    assert 50 < len(sp) < 2000
    assert all(filter_config(c, 2, 1) for c in sp)
    # tiles that do not cover the halo are dropped (reference tuning.py:27)
    assert not filter_config(Config(step=40, bx=64, sn=64, mx=1), 2, 1)
    # fp32 with two vectors per thread exceeds the 256-element TMA box
    assert not filter_config(Config(step=1, bx=64, sn=64, mx=2, dtype="f32"), 2, 2, esize=4)
    # 3D ring must hold the whole k window
    assert not filter_config(Config(step=1, bx=32, by=2, streaming=False, stages=2, rows_3d=8), 3, 1)
    assert search_space(3, 1)


def test_knobs_roundtrip(built):
    import drstencil_b200 as drs
    from drstencil_b200.tuner.space import Config
    c = Config(step=4, bx=64, sn=256, s_unroll=8, mx=2, stages=2, min_blocks=4)
    st = drs.Stencil.from_file(os.path.join(ROOT, "stc", "2d9pt_box.stc")).set_size((1024, 1024))
    info = drs.Plan(st, c.knobs()).info
    assert (info.warps_per_cta, info.chunk, info.stages, info.rows_per_stage, info.tile_x) == (2, 256, 2, 8, 120)


def test_ncu_parser_by_name():
    from drstencil_b200.tuner import metrics
    rows = metrics.parse(open(os.path.join(ROOT, "tests", "golden", "ncu_long_sample.csv")).read())
    assert len(rows) == 5
    dr = [r for r in rows if "dr_c2" in r["kernel"]]
    assert len(dr) == 3 and all(1e-3 < r["gpu__time_duration.sum"] < 2e-3 for r in dr)   # seconds
    raw = metrics.parse(open(os.path.join(ROOT, "tests", "golden", "ncu_raw_sample.csv")).read())
    s = metrics.summarise(raw)
    assert s["kernel"].startswith("dr_c2") and s["launches"] == 2
    assert 4.2e9 < s["dram_bytes"] < 4.6e9                      # ~ the 4.29 GB algorithmic bytes
    assert 3000 < s["dram_gbs"] < 3600
    assert s["launch__registers_per_thread"] == 130


def test_every_baseline_preset_is_a_tuner_winner_with_its_ncu_evidence():
    """SURVEY 8f-1 / BASELINE north_star: every chosen configuration is justified by ncu-measured DRAM GB/s and L2 /
    shared-memory traffic.  Each BASELINE preset must be the winner of its committed tuner record
    (profiles/r02_tune_<workload>.json: sustained timing, Nsight Compute by metric name)."""
    import json
    from drstencil_b200.presets import TUNED
    for wl, (path, dim, name) in TUNED.items():
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_tune_%s.json" % wl)))
        assert rec["stencil"] == os.path.basename(path) and rec["min_seconds_per_candidate"] >= 0.5
        w = rec["winners"][0]
        assert w["name"] == name, (wl, w["name"], name)
        assert w["ms_confirmed"] <= min(x["ms_confirmed"] for x in rec["winners"]) + 1e-12
        assert w["sweeps_confirmed"] * w["ms_confirmed"] >= 900          # >= 0.9 s of back-to-back sweeps
        ncu = w["ncu"]
        assert "error" not in ncu and ncu["launches"] >= 2
        for key in ("dram_gbs", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
                    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__time_duration.sum"):
            assert ncu.get(key, 0) > 0, (wl, key)
        assert 3000 < ncu["dram_gbs"] < 8000
        # the record holds the whole search: the space, a burst figure for every configuration, sustained figures
        # for the finalists
        assert rec["tried"] == rec["space"] == len(rec["all"])
        assert sum(1 for r in rec["all"] if r.get("ms") is not None) >= rec["space"] // 4
