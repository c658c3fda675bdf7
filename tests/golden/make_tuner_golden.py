#!/usr/bin/env python
"""Generates tests/golden/ref_tuner.json by IMPORTING the reference's own tuner scripts
(/root/reference/benchmarks/<stencil>/tuning.py -- plain Python, importable in the authoring container) and
recording, for seeded configuration vectors of every shipped benchmark directory, what the reference itself answers:

  FilterParams(v)       is the configuration admitted to the search            (tuning.py:13-48 / 3d: 13-36)
  cfgToString(v)        the result-file name                                   (tuning.py:72-86 / 3d: 57-72)
  cfgToCommandLine(v)   the `drstencil` arguments it is generated with         (tuning.py:50-69 / 3d: 38-55)

plus the module's `order` (the stencil radius the filter uses).  Half of the vectors are drawn from the whole
product (mostly rejected), half from the admitted ones, steps 1..4.  tests/test_tuner_ref_golden.py holds the
restatements (oracle/tune_ref.py, drstencil_b200/tuner/space.py) to these answers.

Usage (authoring container only; the GPU box has no /root/reference):
    python tests/golden/make_tuner_golden.py
"""
import importlib.util
import itertools
import json
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/benchmarks"
STENCILS = ["2d5pt_star", "2d5pt_cross", "2d9pt_star", "2d9pt_box", "2d9pt_cross", "2d25pt_box", "3d7pt_star", "3d9pt_cross"]
PER = 60            # vectors per benchmark and half


def load(stem):
    spec = importlib.util.spec_from_file_location("ref_tuning_" + stem, os.path.join(REF, stem, "tuning.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # searchSpace() runs only under __main__
    return mod


def vectors(is3d):
    pow2 = [2 ** i for i in range(0, 10)]
    common = ([8, 16, 32, 64], [4, 8], [False, True], [1, 2, 4], [False, True], [1, 2, 4], [1, 5, 9], [False, True])
    blocks = list(itertools.product(pow2, repeat=2))
    if is3d:
        return itertools.product(range(1, 5), range(0, 6), blocks, *common)
    return itertools.product(range(1, 5), range(0, 9), blocks, [False, True], *common)


def main():
    rng = random.Random(20260218)
    out = {}
    for stem in STENCILS:
        mod = load(stem)
        allv = list(vectors(stem.startswith("3d")))
        picked = rng.sample(allv, PER)
        admitted = [v for v in allv if mod.FilterParams(v)]
        picked += rng.sample(admitted, PER)
        rows = []
        for v in picked:
            rows.append({"v": [list(x) if isinstance(x, tuple) else x for x in v], "admit": bool(mod.FilterParams(v)),
                         "name": mod.cfgToString(v), "cmd": mod.cfgToCommandLine(v)})
        out[stem] = {"order": mod.order, "shm_lg2": mod.maxShmPerBlockLg2, "threads_lg2": mod.maxThreadsPerBlockLg2,
                     "admitted_total": len(admitted), "rows": rows}
    path = os.path.join(ROOT, "tests", "golden", "ref_tuner.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d benchmarks, %d vectors" % (path, len(out), sum(len(b["rows"]) for b in out.values())))


if __name__ == "__main__":
    main()
