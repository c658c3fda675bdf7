"""The tuner's search-space filter, result-file names and generator command lines against what the reference's own
tuning.py scripts answer (tests/golden/ref_tuner.json, made by tests/golden/make_tuner_golden.py importing
/root/reference/benchmarks/*/tuning.py).  Two restatements are held to it: oracle/tune_ref.py (which builds the
reference's candidates for the tuned-reference comparison) and drstencil_b200/tuner/space.py (the engine's tuner,
whose names must stay readable by scripts that key on the reference grammar)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import tune_ref  # noqa: E402

from drstencil_b200.tuner import space  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "ref_tuner.json")) as f:
    GOLD = json.load(f)


def vec(row):
    return tuple(tuple(x) if isinstance(x, list) else x for x in row["v"])


def strip_step(tokens):
    i = tokens.index("--step")
    return tokens[:i] + tokens[i + 2:]


@pytest.mark.parametrize("stem", sorted(GOLD))
def test_reference_candidates_restated(stem):
    """oracle/tune_ref.py: same admitted set, same names, same drstencil arguments (--step is added by the caller)."""
    b = GOLD[stem]
    assert (b["threads_lg2"], b["shm_lg2"]) == (tune_ref.MAX_THREADS_LG2, tune_ref.MAX_SHM_LG2)
    is3d = stem.startswith("3d")
    flt, name, cmd = (tune_ref.filter_3d, tune_ref.name_3d, tune_ref.cmdline_3d) if is3d else \
                     (tune_ref.filter_2d, tune_ref.name_2d, tune_ref.cmdline_2d)
    assert sum(r["admit"] for r in b["rows"]) >= 60
    for r in b["rows"]:
        v = vec(r)
        assert flt(v, b["order"]) == r["admit"], (stem, v)
        assert name(v) == r["name"], (stem, v)
        assert sorted(cmd(v)) == sorted(strip_step(r["cmd"].split())), (stem, v)


def as_config(v, is3d):
    if is3d:
        step, dist, bs, sn, unroll, bmx, mx, bmy, my, mf, prefetch = v
        streaming = True
    else:
        step, dist, bs, streaming, sn, unroll, bmx, mx, bmy, my, mf, prefetch = v
    return space.Config(step=step, dist=dist, bx=bs[0], by=bs[1], streaming=streaming, sn=sn, s_unroll=unroll,
                        block_merge_x=bmx, mx=mx, block_merge_y=bmy, my=my, merge_forward=mf, prefetch=prefetch,
                        dim=3 if is3d else 2)


@pytest.mark.parametrize("stem", sorted(GOLD))
def test_engine_tuner_names_follow_the_reference_grammar(stem):
    """drstencil_b200/tuner/space.py: with the engine-only axes at their defaults a configuration carries exactly
    the reference's name, the name parses back to the same configuration, and the command line holds the
    reference's arguments (the engine also states --by and the y merge of a non-streaming 2D configuration, which
    the reference's cfgToCommandLine leaves at the generator defaults -- tuning.py:51-69)."""
    is3d = stem.startswith("3d")
    for r in GOLD[stem]["rows"]:
        v = vec(r)
        c = as_config(v, is3d)
        assert space.cfg_to_string(c) == r["name"], (stem, v)
        back = space.cfg_from_string(r["name"], 3 if is3d else 2)
        assert space.cfg_to_string(back) == r["name"]
        assert (back.step, back.dist, back.bx, back.mx, back.block_merge_x, back.merge_forward) == \
               (c.step, c.dist, c.bx, c.mx, c.block_merge_x, c.merge_forward)
        mine = space.cfg_to_command_line(c).split()
        ref = r["cmd"].split()
        pairs = lambda t: sorted(" ".join(t[i:i + 2]) if i + 1 < len(t) and not t[i + 1].startswith("--") else t[i]
                                 for i in range(len(t)) if t[i].startswith("--"))
        missing = [p for p in pairs(ref) if p not in pairs(mine)]
        assert not missing, (stem, v, missing)
        extra = [p for p in pairs(mine) if p not in pairs(ref)]
        if is3d or c.streaming:
            assert not extra, (stem, v, extra)
        else:
            assert all(p.startswith(("--by ", "--block-merge-y ", "--cyclic-merge-y ")) for p in extra), (stem, v, extra)
