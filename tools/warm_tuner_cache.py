#!/usr/bin/env python
"""Pre-compiles (NVRTC, no GPU) every configuration of the tuner's search spaces for the BASELINE workloads into
drstencil_b200/_jitcache/, in parallel, so that a tuner run on the GPU box spends its time measuring."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(job):
    path, dim, name = job
    import drstencil_b200 as drs
    from drstencil_b200.tuner.space import cfg_from_string
    try:
        drs.Plan(drs.Stencil.from_file(path), cfg_from_string(name, dim).knobs())
        return None
    except Exception as e:
        return "%s: %s" % (name, str(e)[:100])


def main():
    import drstencil_b200 as drs
    from drstencil_b200.presets import TUNED
    from drstencil_b200.tuner.space import cfg_to_string, search_space
    jobs = []
    for wl, (path, dim, preset) in TUNED.items():
        st = drs.Stencil.from_file(path)
        radius = max(max(abs(t[0]), abs(t[1]), abs(t[2])) for t in st.terms())
        from drstencil_b200.tuner.space import cfg_from_string
        c0 = cfg_from_string(preset, dim)
        for c in search_space(dim, radius, c0.step, c0.dtype, c0.fuse):
            jobs.append((path, dim, cfg_to_string(c)))
    print("warm_tuner_cache: %d configurations" % len(jobs))
    with ProcessPoolExecutor(int(os.environ.get("JOBS", "8"))) as ex:
        errs = [e for e in ex.map(one, jobs, chunksize=8) if e]
    print("warm_tuner_cache: done, %d refused by the engine" % len(errs))
    for e in errs[:10]:
        print("  ", e)


if __name__ == "__main__":
    main()
