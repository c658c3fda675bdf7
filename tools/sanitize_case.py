"""Tiny sweeps of every kernel family, for compute-sanitizer (memcheck): 2D gather, 2D scatter
(+ factorised, 2 vectors/thread), 3D, gold, check.  Prints SANITIZE_CASE_OK when results match
the gold kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import drstencil_b200 as drs

ok = True
for name, shape, kn in [("2d5pt_star", (70, 136), dict(sn=16)), ("2d9pt_box", (90, 264), dict(step=4, vectors=2, sn=24)),
                        ("2d25pt_box", (60, 200), dict(dtype="f32", step=2)), ("2d9pt_star", (50, 64), dict(step=2)),
                        ("3d7pt_star", (14, 20, 70), dict(sn=5, rows_3d=4)), ("3d9pt_cross", (12, 18, 66), dict())]:
    st = drs.Stencil.from_file(os.path.join(ROOT, "stc", name + ".stc")).set_size(shape)
    plan = drs.Plan(st, drs.Knobs(**kn))
    dt = torch.float32 if kn.get("dtype") == "f32" else torch.float64
    A = torch.rand(shape, dtype=dt, device="cuda")
    B, G = torch.zeros_like(A), torch.zeros_like(A)
    plan.sweep(A, B)
    plan.gold_sweep(A, G)
    plan.sync_check()
    mx, rms = plan.check_error(B, G)
    good = mx < (1e-4 if dt == torch.float32 else 1e-12)
    print("%-12s %-14s %s max|err| %.3e %s" % (name, shape, kn, mx, "ok" if good else "MISMATCH"), flush=True)
    ok = ok and good
print("SANITIZE_CASE_OK" if ok else "SANITIZE_CASE_FAILED")
sys.exit(0 if ok else 1)
