"""Named configurations: the eight stencil descriptions the reference ships
(/root/reference/benchmarks/*/*.stc, re-typed under stc/) and the five BASELINE.json workloads.

The BASELINE entries are TUNER RECORDS: each is stated as the configuration name (the reference's result-file
grammar, tuning.py:72-86, plus the engine's axes -- tuner/space.py) of the winner of a tuner run on B200, whose
record under profiles/r02_tune_<workload>.json carries the Nsight Compute metrics that justify it (DRAM bytes and
GB/s, L2 and shared-memory traffic).  tests/test_tuner.py checks every entry against its record."""
import os

from . import Knobs
from .tuner.space import cfg_from_string

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STC = os.path.join(_ROOT, "stc")


def _p(*parts):
    return os.path.join(_STC, *parts)


# workload -> (stc file, dimensionality, configuration name == winners[0].name of profiles/r02_tune_<workload>.json)
TUNED = {
    "c1": (_p("baseline", "c1_2d5pt_star.stc"), 2, "fu1d0bx32sn128u4bmx2mf5st2"),
    "c2": (_p("baseline", "c2_2d9pt_box.stc"), 2, "fu4d0bx64sn256u4bmx2mf5st2mb4"),
    "c3": (_p("baseline", "c3_2d25pt_box.stc"), 2, "fu1d0bx64sn32u8bmx1mf5st2f32"),
    # c4 / c5: four warps (2 x 2) share one input ring per CTA (drs_sweep3d_cta.cuh); c5 with eight rows per thread
    # (9.32 ms sustained vs 9.50 for round 1's private rings with six rows)
    "c4": (_p("baseline", "c4_3d7pt_star.stc"), 3, "fu1d0bx32y4sn64u4bmx1bmy1mf5ry4sx2sy2"),
    "c5": (_p("baseline", "c5_3d7pt_star.stc"), 3, "fu1d0bx32y4sn64u4bmx1bmy1mf5ry8sx2sy2"),
}


def _tuned(name):
    path, dim, cfg = TUNED[name]
    return (path, cfg_from_string(cfg, dim).knobs())


# name -> (stc file, knobs)
PRESETS = {
    "2d5pt_star": (_p("2d5pt_star.stc"), Knobs()),
    "2d5pt_cross": (_p("2d5pt_cross.stc"), Knobs()),
    "2d9pt_star": (_p("2d9pt_star.stc"), Knobs()),
    "2d9pt_box": (_p("2d9pt_box.stc"), Knobs()),
    "2d9pt_cross": (_p("2d9pt_cross.stc"), Knobs()),
    "2d25pt_box": (_p("2d25pt_box.stc"), Knobs()),
    "3d7pt_star": (_p("3d7pt_star.stc"), Knobs()),
    "3d9pt_cross": (_p("3d9pt_cross.stc"), Knobs()),
    # BASELINE.json configs (tuner records, see TUNED)
    "c1": _tuned("c1"), "c2": _tuned("c2"), "c3": _tuned("c3"), "c4": _tuned("c4"), "c5": _tuned("c5"),
    # extra (not a BASELINE.json config): c4 with in-kernel temporal depth 2
    "c4t2": (_p("baseline", "c4_3d7pt_star.stc"), Knobs(step=2)),
    "c5t2": (_p("baseline", "c5_3d7pt_star.stc"), Knobs(step=2)),
    "c1t2": (_p("baseline", "c1_2d5pt_star.stc"), Knobs(step=2)),
    "c3t2": (_p("baseline", "c3_2d25pt_box.stc"), Knobs(dtype="f32", step=2, sn=256)),
}
