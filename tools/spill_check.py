"""ptxas resource lines (registers, spills) of the sweep entry points of the named presets (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drstencil_b200 as drs, os, sys
from drstencil_b200.presets import PRESETS
drs.build()
for name in sys.argv[1:]:
    path, kn = PRESETS[name]
    plan = drs.Plan(drs.Stencil.from_file(path), kn)
    log = open(os.path.join(os.path.dirname(drs.__file__), "_jitcache", plan.cache_key + ".log")).read()
    for entry in ("dr_", "drslab_"):
        if "entry function '" + entry not in log:
            continue
        i = log.index("entry function '" + entry)
        print(name, entry, " | ".join(l.strip() for l in log[i:].splitlines()[2:4]))
