#!/usr/bin/env python
"""Sustained timing (tuner.time_config, >= 1 s of back-to-back sweeps) of named presets -- development aid.
usage: python tools/time_presets.py c2 c3t2 ...   (environment switches such as DRS_NO_DEFERRED_SCALE apply)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    from drstencil_b200.tuner.space import Config
    from drstencil_b200.tuner import tune

    class K:        # time_config wants an object with knobs(), step, dtype
        def __init__(self, kn):
            self.kn, self.step, self.dtype = kn, kn.step, "f32" if kn.dtype == drs.F32 else "f64"

        def knobs(self):
            return self.kn

    for name in sys.argv[1:]:
        path, kn = PRESETS[name]
        st = drs.Stencil.from_file(path)
        ms, info, n = tune.time_config(st, K(kn), min_seconds=float(os.environ.get("MIN_SECONDS", "1.0")), warm=4)
        pts = 1
        allp = 1
        for d in st.shape:
            pts *= d - 2 * info.halo
            allp *= d
        es = 4 if kn.dtype == drs.F32 else 8
        gbs = allp * 2 * es / (ms * 1e-3) / 1e9
        print("%-6s %.4f ms per sweep over %d sweeps  %.1f GStencil/s  %.0f GB/s (%.1f %% of 6553.6)  regs %d"
              % (name, ms, n, pts * kn.step / (ms * 1e-3) / 1e9, gbs, 100 * gbs / 6553.6, info.regs_per_thread), flush=True)


if __name__ == "__main__":
    main()
