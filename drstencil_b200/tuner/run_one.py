"""Launches one configuration a few times on device-resident data: the process Nsight Compute
wraps (stand-in for the reference's compiled bin/<cfg>, compile_run.sh:5).

    python -m drstencil_b200.tuner.run_one <stc> [--3d] [--size L M N] [--launches n] -- <drstencil options>
    python -m drstencil_b200.tuner.run_one --preset c2 [--launches n]
"""
import argparse
import sys

import torch

from .. import Knobs, Plan, Stencil, F32


def knobs_from_argv(argv):
    k = Knobs()
    i = 0
    flags = {"--streaming": "streaming", "--prefetch": "prefetch", "--check": "check"}
    while i < len(argv):
        a = argv[i]
        if a in flags:
            k.set(flags[a], 1)
            i += 1
        elif a.startswith("--") and i + 1 < len(argv):
            name = a[2:].replace("-", "_")
            k.set(name, argv[i + 1] if name in ("dtype", "fuse") else int(argv[i + 1]))
            i += 2
        else:
            raise SystemExit("bad option %r" % a)
    return k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stc", nargs="?")
    ap.add_argument("--preset", help="a name from drstencil_b200/presets.py instead of <stc> + options")
    ap.add_argument("--3d", dest="is3d", action="store_true")
    ap.add_argument("--size", type=int, nargs="+")
    ap.add_argument("--launches", type=int, default=6)
    argv = sys.argv[1:]
    rest = []
    if "--" in argv:
        cut = argv.index("--")
        argv, rest = argv[:cut], argv[cut + 1:]
    a = ap.parse_args(argv)
    if a.preset:
        from ..presets import PRESETS
        a.stc, kn = PRESETS[a.preset]
    else:
        kn = knobs_from_argv(rest)
    st = Stencil.from_file(a.stc, a.is3d or None)
    if a.size:
        st.set_size(a.size)
    plan = Plan(st, kn)
    dt = torch.float32 if kn.dtype == F32 else torch.float64
    A = torch.rand(st.shape, dtype=dt, device="cuda")
    B = torch.zeros_like(A)
    bufs = [A, B]
    for s in range(a.launches):
        plan.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
    plan.sync_check()
    print("run_one: %s %d launches ok" % (plan.info.kernel_name, a.launches))


if __name__ == "__main__":
    main()
