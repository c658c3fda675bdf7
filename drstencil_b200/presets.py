"""Named configurations: the eight stencil descriptions the reference ships
(/root/reference/benchmarks/*/*.stc, re-typed under stc/) and the five BASELINE.json workloads.
Knob values for the BASELINE entries are what the tuner (drstencil_b200/tuner) settled on; the
evidence is under profiles/."""
import os

from . import Knobs

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STC = os.path.join(_ROOT, "stc")


def _p(*parts):
    return os.path.join(_STC, *parts)


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
    # BASELINE.json configs
    "c1": (_p("baseline", "c1_2d5pt_star.stc"), Knobs(sn=128, warps=2, vectors=2, stages=2, rows_per_stage=4)),
    "c2": (_p("baseline", "c2_2d9pt_box.stc"), Knobs(step=4, sn=256, vectors=2, stages=2)),
    "c3": (_p("baseline", "c3_2d25pt_box.stc"), Knobs(dtype="f32", sn=32, warps=2, rows_per_stage=8, stages=2, min_blocks=4)),
    "c4": (_p("baseline", "c4_3d7pt_star.stc"), Knobs(sn=16, rows_3d=4)),
    # chunk 64 is best under sustained load (power cap); four x-adjacent warps per CTA read 4 % less from DRAM than two
    "c5": (_p("baseline", "c5_3d7pt_star.stc"), Knobs(sn=64, rows_3d=6, warps=4)),
    # extra (not a BASELINE.json config): c4 with in-kernel temporal depth 2
    "c4t2": (_p("baseline", "c4_3d7pt_star.stc"), Knobs(step=2)),
    "c5t2": (_p("baseline", "c5_3d7pt_star.stc"), Knobs(step=2)),
    "c1t2": (_p("baseline", "c1_2d5pt_star.stc"), Knobs(step=2)),
    "c3t2": (_p("baseline", "c3_2d25pt_box.stc"), Knobs(dtype="f32", step=2, sn=256)),
}
