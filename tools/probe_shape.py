#!/usr/bin/env python
"""Event timing of one shipped stencil at an arbitrary grid size for a list of knob sets.
usage: python tools/probe_shape.py <stencil> <L,M,N | M,N> '<json knobs>' ['<json knobs>' ...]
Development aid (like tools/probe.py), not the contract bench."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import drstencil_b200 as drs
    from probe import time_plan, report
    name, shape = sys.argv[1], tuple(int(x) for x in sys.argv[2].split(","))
    for js in sys.argv[3:]:
        kn = json.loads(js)
        st = drs.Stencil.from_file(os.path.join(ROOT, "stc", name + ".stc")).set_size(shape)
        try:
            plan = drs.Plan(st, drs.Knobs(**kn))
            dtype = torch.float32 if kn.get("dtype") == "f32" else torch.float64
            best = min(time_plan(plan, shape, dtype) for _ in range(3))
            report("%s %s" % (name, "x".join(map(str, shape))), plan, shape, dtype, best)
        except Exception as e:
            print(name, kn, "FAILED:", str(e)[:300], flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
